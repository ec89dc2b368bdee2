__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__global__ void k(const float2* a, const float2* b, float2* c) {
  float2 x = a[threadIdx.x], y = b[threadIdx.x];
  unsigned long long X = pk(x.x, x.y), Y = pk(y.x, y.y);
  unsigned long long acc = X;
  for (int i = 0; i < 4; ++i) acc = add2(acc, mul2(X, Y));
  float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
  c[threadIdx.x] = make_float2(lo, hi);
}
