// Bring-up harness (NOT part of the product library): the per-image moments sweep of image_normalize on the batch-fastest
// layout x[b + B*p] (f32, B = 8 images of 2160 x 3840 pixels = 265 MB), DESIGN.md section 8 item 3.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o moments_variants moments_variants.cu && ./moments_variants
//
// r04: the product kernel (variant 0 below) takes 93.5 us = 2.84 TB/s; ncu: issue slots 26 % busy, stall_short_scoreboard 10.8,
// stall_barrier 3.3 - every thread ends with 8 shared-memory f64 atomics onto 16 addresses (128-way contention, and an
// arrival-order-dependent sum). Variant 1 folds the lanes that share an image with warp shuffles, writes per-warp slots and
// folds them in a fixed order: no atomics, bit-reproducible. Variant 2 adds 16 loads in flight.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int B = 8, VEC = 4;

__host__ __device__ inline float pix(u64 i) {  // the benchmark's LCG field (4k-image-processing/runmat_lcg.m:59-79)
  const u32 s = (u32)(1664525ull * i + 1013904223ull);
  return (float)((double)s / 4294967296.0);
}
__global__ void fill_kernel(float* x, u64 n) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) x[i] = pix(i);
}

// variant 0: the product kernel as of r04 (shifted sums around K[b] = x[b], shared-memory atomics at the end)
template <int U>
__global__ void __launch_bounds__(256) moments_atomic(const float* __restrict__ x, u64 total, double* __restrict__ partial /*[grid][2][B]*/) {
  __shared__ double sh[2 * B];
  if (threadIdx.x < 2 * B) sh[threadIdx.x] = 0.0;
  __syncthreads();
  const u64 nvec = total / VEC, nthr = (u64)gridDim.x * blockDim.x;
  const u64 v0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 b0 = (u32)((v0 * VEC) % B);
  double K[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { K[l] = (double)x[b0 + l]; s1[l] = 0.0; s2[l] = 0.0; }
  const float4* xv = reinterpret_cast<const float4*>(x);
  u64 v = v0;
  for (; v + (u64)(U - 1) * nthr < nvec; v += (u64)U * nthr) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = __ldcs(xv + v + (u64)u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* e = reinterpret_cast<const float*>(&a[u]);
#pragma unroll
      for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
    }
  }
  for (; v < nvec; v += nthr) {
    const float4 a = __ldcs(xv + v);
    const float* e = reinterpret_cast<const float*>(&a);
#pragma unroll
    for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
  }
#pragma unroll
  for (int l = 0; l < VEC; ++l) { atomicAdd(&sh[b0 + l], s1[l]); atomicAdd(&sh[B + b0 + l], s2[l]); }
  __syncthreads();
  if (threadIdx.x < 2 * B) partial[(u64)blockIdx.x * 2 * B + threadIdx.x] = sh[threadIdx.x];
}

// variants 1/2: deterministic fold. Lanes of a warp alternate between images {0..3} (even lane) and {4..7} (odd lane) because
// consecutive threads own consecutive float4's and B / VEC = 2; xor-shuffles over offsets 2..16 keep the parity.
template <int U>
__global__ void __launch_bounds__(256) moments_fold(const float* __restrict__ x, u64 total, double* __restrict__ partial /*[grid][2][B]*/) {
  __shared__ double shw[8][2][B];  // [warp][moment][image]
  const u64 nvec = total / VEC, nthr = (u64)gridDim.x * blockDim.x;
  const u64 v0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u32 b0 = (u32)((v0 * VEC) % B);
  double K[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { K[l] = (double)x[b0 + l]; s1[l] = 0.0; s2[l] = 0.0; }
  const float4* xv = reinterpret_cast<const float4*>(x);
  u64 v = v0;
  for (; v + (u64)(U - 1) * nthr < nvec; v += (u64)U * nthr) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = __ldcs(xv + v + (u64)u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* e = reinterpret_cast<const float*>(&a[u]);
#pragma unroll
      for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
    }
  }
  for (; v < nvec; v += nthr) {
    const float4 a = __ldcs(xv + v);
    const float* e = reinterpret_cast<const float*>(&a);
#pragma unroll
    for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < VEC; ++l)
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) {
      s1[l] += __shfl_xor_sync(0xffffffffu, s1[l], off);
      s2[l] += __shfl_xor_sync(0xffffffffu, s2[l], off);
    }
  if (lane < 2) {  // lane 0 holds images b0..b0+3 of the even lanes, lane 1 those of the odd lanes
#pragma unroll
    for (int l = 0; l < VEC; ++l) { shw[warp][0][b0 + l] = s1[l]; shw[warp][1][b0 + l] = s2[l]; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * B) {
    const int m = threadIdx.x / B, b = threadIdx.x % B;
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += shw[w][m][b];  // warp order: deterministic
    partial[(u64)blockIdx.x * 2 * B + threadIdx.x] = t;
  }
}

int main() {
  const u64 P = 2160ull * 3840ull, total = P * B;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* x;
  double* partial;
  const int grid = sms * 8;  // (grid * 256 * VEC) % B == 0 keeps every thread on a fixed image group
  CK(cudaMalloc(&x, total * 4));
  CK(cudaMalloc(&partial, (size_t)grid * 2 * B * 8));
  fill_kernel<<<sms * 8, 256>>>(x, total);
  CK(cudaDeviceSynchronize());
  // host reference: per-image mean and population variance in f64
  std::vector<double> m(B, 0.0), q(B, 0.0);
  for (u64 i = 0; i < total; ++i) m[i % B] += (double)pix(i);
  for (int b = 0; b < B; ++b) m[b] /= (double)P;
  for (u64 i = 0; i < total; ++i) { const double d = (double)pix(i) - m[i % B]; q[i % B] += d * d; }

  struct V { const char* name; void (*run)(const float*, u64, double*, int); };
  const V variants[] = {
      {"atomic U=8 (product r04)", [](const float* x, u64 t, double* p, int g) { moments_atomic<8><<<g, 256>>>(x, t, p); }},
      {"fold   U=8 (deterministic)", [](const float* x, u64 t, double* p, int g) { moments_fold<8><<<g, 256>>>(x, t, p); }},
      {"fold   U=16", [](const float* x, u64 t, double* p, int g) { moments_fold<16><<<g, 256>>>(x, t, p); }},
      {"fold   U=4", [](const float* x, u64 t, double* p, int g) { moments_fold<4><<<g, 256>>>(x, t, p); }},
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<double> hp((size_t)grid * 2 * B);
  for (const V& v : variants) {
    for (int w = 0; w < 3; ++w) v.run(x, total, partial, grid);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 50; ++r) v.run(x, total, partial, grid);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(hp.data(), partial, hp.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int b = 0; b < B; ++b) {
      double s1 = 0.0, s2 = 0.0;
      for (int g = 0; g < grid; ++g) { s1 += hp[(size_t)g * 2 * B + b]; s2 += hp[(size_t)g * 2 * B + B + b]; }
      const double K = (double)pix((u64)b), dm = s1 / (double)P, mean = K + dm, var = (s2 - s1 * dm) / (double)P;
      worst = fmax(worst, fmax(fabs(mean - m[b]) / fabs(m[b]), fabs(var - q[b] / (double)P) / (q[b] / (double)P)));
    }
    const double us = ms * 1e3 / 50;
    printf("%-30s %7.1f us  %6.0f GB/s  worst rel.err(mean,var) %.2e\n", v.name, us, total * 4.0 / us / 1e3, worst);
  }
  return 0;
}
