// Bring-up harness (NOT part of the product library): variants of the headline fused reduction
//     s = sum(sin(A) .* B + 1),  A, B in f64[4096 x 4096]   (16 B/elem algorithmic, 268.4 MB per launch)
// timed side by side in ONE run, so round 2 can pick the structure for fusion_lower.cpp's Contig emitter with a single GPU
// call. Every variant keeps the product's contract: 256-bit L1-no-allocate loads, f64
// accumulation, deterministic two-stage finish (last block by ticket), no floating-point atomics, -fmad=false.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o reduction_variants reduction_variants.cu
//   ./reduction_variants [n_side=4096] [reps=50]
//
// r04 baseline to beat: 53.5 us = 5.02 TB/s = 0.775 of the measured copy bandwidth; ncu: issue slots 61 % busy at 46 % warps
// active (63 regs), stall_wait 2.9 (dependent DADD chain of the single accumulator), long_scoreboard 4.0.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
typedef unsigned long long u64;
typedef unsigned int u32;
struct __align__(32) vec_t { double x[4]; };
__device__ __forceinline__ vec_t ldv(const double* p) {
  vec_t r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x[0]), "=d"(r.x[1]), "=d"(r.x[2]), "=d"(r.x[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ double expr(double a, double b) { return sin(a) * b + 1.0; }

// ---- data: a counter-based hash so host and device agree without an upload ------------------------------------------------
__host__ __device__ inline double u01(u64 i, u64 salt) {
  u64 z = i * 0x9E3779B97F4A7C15ull + salt;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void fill_kernel(double* A, double* B, u64 n) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    A[i] = u01(i, 1) * 12.566370614359172;  // U(0, 4 pi)
    B[i] = u01(i, 2) * 2.0 - 1.0;           // U(-1, 1)
  }
}

// ---- shared pieces ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) { r = lane < THREADS / 32 ? smem[lane] : 0.0; r = warp_sum(r); }
  return r;
}
template <int THREADS>
__device__ __forceinline__ void finish(double total, double* partial, u32* ticket, double* out, double* smem, bool* is_last) {
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = total;
    __threadfence();
    *is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!*is_last) return;
  __threadfence();
  double a = 0.0;
  for (u32 b = threadIdx.x; b < gridDim.x; b += THREADS) a += __ldcg(&partial[b]);
  a = block_sum<THREADS>(a, smem);
  if (threadIdx.x == 0) { *out = a; *ticket = 0; }
}

// NACC independent accumulators (element l of a vector goes to accumulator l % NACC); TRACK = the product's per-element NaN
// bookkeeping; U = vectors per input in flight per iteration.
template <int THREADS, int MINB, int NACC, bool TRACK, int U>
__global__ void __launch_bounds__(THREADS, MINB)
red_kernel(const double* __restrict__ A, const double* __restrict__ B, u64 n, double* __restrict__ partial, u32* __restrict__ ticket,
           double* __restrict__ out) {
  __shared__ double smem[32];
  __shared__ bool is_last;
  const u64 nvec = n / 4, nthr = (u64)gridDim.x * THREADS;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  bool saw_nan = false;
  u64 i = (u64)blockIdx.x * THREADS + threadIdx.x;
  for (; i + (u64)(U - 1) * nthr < nvec; i += (u64)U * nthr) {
    vec_t a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a[u] = ldv(A + (i + (u64)u * nthr) * 4); b[u] = ldv(B + (i + (u64)u * nthr) * 4); }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const double v = expr(a[u].x[l], b[u].x[l]);
        if (TRACK) { if (v != v) saw_nan = true; else acc[(u * 4 + l) % NACC] += v; }
        else acc[(u * 4 + l) % NACC] += v;
      }
  }
  for (; i < nvec; i += nthr) {
    const vec_t a = ldv(A + i * 4), b = ldv(B + i * 4);
#pragma unroll
    for (int l = 0; l < 4; ++l) acc[l % NACC] += expr(a.x[l], b.x[l]);
  }
  double t = acc[0];
#pragma unroll
  for (int k = 1; k < NACC; ++k) t += acc[k];  // fixed fold order
  if (TRACK && __syncthreads_or(saw_nan ? 1 : 0)) t = __longlong_as_double(0x7ff8000000000000LL);
  t = block_sum<THREADS>(t, smem);
  finish<THREADS>(t, partial, ticket, out, smem, &is_last);
}

struct Variant { const char* name; const void* fn; int threads; void (*launch)(int, const double*, const double*, u64, double*, u32*, double*, cudaStream_t); };
template <int THREADS, int MINB, int NACC, bool TRACK, int U>
void launch_v(int grid, const double* A, const double* B, u64 n, double* partial, u32* ticket, double* out, cudaStream_t st) {
  red_kernel<THREADS, MINB, NACC, TRACK, U><<<grid, THREADS, 0, st>>>(A, B, n, partial, ticket, out);
}
#define V(name, T, M, N, K, U) {name, (const void*)red_kernel<T, M, N, K, U>, T, launch_v<T, M, N, K, U>}

int main(int argc, char** argv) {
  const u64 side = argc > 1 ? strtoull(argv[1], nullptr, 10) : 4096;
  const int reps = argc > 2 ? atoi(argv[2]) : 50;
  const u64 n = side * side;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *A, *B, *partial, *out;
  u32* ticket;
  CK(cudaMalloc(&A, n * 8)); CK(cudaMalloc(&B, n * 8));
  CK(cudaMalloc(&partial, 65536 * 8)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&ticket, 4));
  CK(cudaMemset(ticket, 0, 4));
  fill_kernel<<<sms * 8, 256>>>(A, B, n);
  CK(cudaDeviceSynchronize());
  // reference: Kahan-free plain f64 sum on the host over the same hash (glibc sin); parity bar 1e-10 * sum|terms|
  double ref = 0.0, refabs = 0.0;
  for (u64 i = 0; i < n; ++i) { const double v = std::sin(u01(i, 1) * 12.566370614359172) * (u01(i, 2) * 2.0 - 1.0) + 1.0; ref += v; refabs += std::fabs(v); }

  const Variant variants[] = {
      V("baseline   256thr x4/SM 1acc track U2", 256, 1, 1, true, 2),
      V("no-track   256thr x4/SM 1acc       U2", 256, 1, 1, false, 2),
      V("2 accs     256thr x4/SM            U2", 256, 1, 2, false, 2),
      V("4 accs     256thr x4/SM            U2", 256, 1, 4, false, 2),
      V("2 accs     256thr x4/SM minb4      U2", 256, 4, 2, false, 2),
      V("2 accs     512thr x2/SM            U2", 512, 1, 2, false, 2),
      V("2 accs     128thr x8/SM            U2", 128, 1, 2, false, 2),
      V("2 accs     256thr x4/SM            U1", 256, 1, 2, false, 1),
      V("4 accs     256thr x3/SM            U4", 256, 1, 4, false, 4),
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("n = %llu elements (%.1f MB algorithmic per launch), %d SMs, reference sum %.15g\n", n, n * 16.0 / 1e6, sms, ref);
  for (const Variant& v : variants) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, v.fn));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, v.threads, 0));
    for (int ctas_per_sm = 1; ctas_per_sm <= per_sm; ctas_per_sm *= 2) {
      const int grid = sms * (ctas_per_sm * 2 > per_sm && ctas_per_sm != per_sm ? per_sm : ctas_per_sm);
      for (int w = 0; w < 3; ++w) v.launch(grid, A, B, n, partial, ticket, out, 0);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; ++r) v.launch(grid, A, B, n, partial, ticket, out, 0);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      double got = 0;
      CK(cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost));
      const double us = ms * 1e3 / reps;
      printf("%-42s regs %3d  grid %5d (%d/SM of %d)  %7.1f us  %6.0f GB/s  rel.err %.2e %s\n", v.name, fa.numRegs, grid, grid / sms, per_sm, us,
             n * 16.0 / us / 1e3, std::fabs(got - ref) / refabs, std::fabs(got - ref) <= 1e-10 * refabs ? "" : "PARITY FAIL");
      if (grid == sms * per_sm) break;
    }
  }
  return 0;
}
