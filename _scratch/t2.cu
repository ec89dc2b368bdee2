__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void k(const unsigned long long* a, const unsigned long long* b, unsigned long long* c, unsigned long long* d) {
  c[threadIdx.x] = mul2(a[threadIdx.x], b[threadIdx.x]);
  d[threadIdx.x] = add2(a[threadIdx.x], b[threadIdx.x]);
}
__global__ void k2(const float* a, const float* b, float* c) {
  float x = a[threadIdx.x], y = b[threadIdx.x], m, s;
  asm("mul.rn.f32 %0, %1, %2;" : "=f"(m) : "f"(x), "f"(y));
  asm("add.rn.f32 %0, %1, %2;" : "=f"(s) : "f"(m), "f"(x));
  c[threadIdx.x] = s;
}
