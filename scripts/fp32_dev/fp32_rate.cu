// Bring-up probe (NOT part of the product): issue cost of the FP32 forms the bit-exact imfilter can use on sm_100a.
// Each kernel runs ITERS x 16 independent ops per thread on 1024-thread CTAs (8 warps per SM sub-partition), one CTA per SM;
// reported: SM cycles per warp instruction per sub-partition (1.0 = one instruction issued every cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
typedef unsigned long long u64;
constexpr int ITERS = 4096, NACC = 16;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
template <int MODE>
__global__ void __launch_bounds__(1024) rate_kernel(float* out, float seed, long long* cycles) {
  float a[NACC]; u64 p[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = seed + i + threadIdx.x; p[i] = pk(a[i], a[i] + 1.f); }
  const float w = seed * 0.5f; const u64 w2 = pk(w, w), one2 = pk(seed * 0.f + 1.f, seed * 0.f + 1.f), c2 = pk(seed, seed);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) a[i] = __fmaf_rn(a[i], w, seed);                                            // scalar FFMA
      if (MODE == 1) a[i] = __fmul_rn(a[i], w);                                                 // scalar FMUL
      if (MODE == 2) a[i] = __fadd_rn(a[i], w);                                                 // scalar FADD
      if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(w2), "l"(c2));      // packed FFMA2, all operands packed registers
      if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(w2));                   // packed FMUL2
      if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(w2));                   // packed FADD2
      if (MODE == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(one2), "l"(w2));    // the imfilter's times-one add
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaMalloc(&out, (size_t)sms * 1024 * 4)); CK(cudaMalloc(&cyc, 8));
  const char* names[] = {"scalar FFMA", "scalar FMUL", "scalar FADD", "packed FFMA2 (all packed)", "packed FMUL2", "packed FADD2", "packed FFMA2 x (1,1) + p"};
  for (int m = 0; m < 7; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (m) {
        case 0: rate_kernel<0><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 1: rate_kernel<1><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 2: rate_kernel<2><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 3: rate_kernel<3><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 4: rate_kernel<4><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 5: rate_kernel<5><<<sms, 1024>>>(out, 1.0f, cyc); break;
        case 6: rate_kernel<6><<<sms, 1024>>>(out, 1.0f, cyc); break;
      }
      CK(cudaDeviceSynchronize());
    }
    long long h = 0; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double per = (double)h / ((double)ITERS * NACC * 8);  // 8 warps per sub-partition
    printf("%-28s %8.3f cycles per warp instruction per sub-partition  (%5.1f results/clk/SM)\n", names[m], per, (m >= 3 ? 64.0 : 32.0) * 4 / per);
  }
  return 0;
}
