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


// ---- round-2 candidates --------------------------------------------------------------------------------------------------
// (a) register double buffering: the loads of iteration i+1 are issued before the math of iteration i (U = 1 vector per input
//     per iteration, so the register footprint equals the U2 kernel's) -- a warp always has loads in flight while it computes.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
red_prefetch_kernel(const double* __restrict__ A, const double* __restrict__ B, u64 n, double* __restrict__ partial, u32* __restrict__ ticket,
                    double* __restrict__ out) {
  __shared__ double smem[32];
  __shared__ bool is_last;
  const u64 nvec = n / 4, nthr = (u64)gridDim.x * THREADS;
  double acc0 = 0.0, acc1 = 0.0;
  u64 i = (u64)blockIdx.x * THREADS + threadIdx.x;
  if (i < nvec) {
    vec_t a = ldv(A + i * 4), b = ldv(B + i * 4);
    for (i += nthr; i < nvec; i += nthr) {
      const vec_t an = ldv(A + i * 4), bn = ldv(B + i * 4);
      acc0 += expr(a.x[0], b.x[0]); acc1 += expr(a.x[1], b.x[1]); acc0 += expr(a.x[2], b.x[2]); acc1 += expr(a.x[3], b.x[3]);
      a = an; b = bn;
    }
    acc0 += expr(a.x[0], b.x[0]); acc1 += expr(a.x[1], b.x[1]); acc0 += expr(a.x[2], b.x[2]); acc1 += expr(a.x[3], b.x[3]);
  }
  double t = block_sum<THREADS>(acc0 + acc1, smem);
  finish<THREADS>(t, partial, ticket, out, smem, &is_last);
}

// (b) TMA bulk-copy ring (north_star: "cp.async/TMA staging into shared memory"): a producer warp keeps STAGES chunks of both
//     inputs in flight with cp.async.bulk (SASS UBLKCP) completing on mbarriers; the compute warps read 128-bit pieces from
//     shared memory (conflict-free: piece j of thread t sits at (j*THREADS + t)*16 B). Bytes in flight no longer depend on the
//     compute warps' phase or on registers.
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(u64* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
struct __align__(16) v2_t { double x, y; };
template <int THREADS, int MINB, int STAGES, int VPT>
__global__ void __launch_bounds__(THREADS + 32, MINB)
red_bulk_kernel(const double* __restrict__ A, const double* __restrict__ B, u64 n, double* __restrict__ partial, u32* __restrict__ ticket,
                double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ double smem[32];
  __shared__ bool is_last;
  __shared__ __align__(8) u64 full[STAGES], empty[STAGES];
  constexpr u32 CHUNK_B = THREADS * VPT * 32;        // bytes per input per stage
  constexpr u64 CHUNK_E = (u64)THREADS * VPT * 4;    // elements per chunk
  const u64 nchunks = n / CHUNK_E;                   // the harness sizes are multiples of the chunk (the product adds a direct tail)
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc0 = 0.0, acc1 = 0.0;
  if (tid >= THREADS) {  // producer warp
    if (tid == THREADS) {
      u32 k = 0;
      for (u64 c = blockIdx.x; c < nchunks; c += gridDim.x, ++k) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty[s], ((k / STAGES) - 1) & 1);
        mbar_expect_tx(&full[s], 2 * CHUNK_B);
        bulk_g2s(dsm + (size_t)s * 2 * CHUNK_B, A + c * CHUNK_E, CHUNK_B, &full[s]);
        bulk_g2s(dsm + (size_t)s * 2 * CHUNK_B + CHUNK_B, B + c * CHUNK_E, CHUNK_B, &full[s]);
      }
    }
  } else {
    u32 k = 0;
    for (u64 c = blockIdx.x; c < nchunks; c += gridDim.x, ++k) {
      const int s = k % STAGES;
      mbar_wait(&full[s], (k / STAGES) & 1);
      const v2_t* a = (const v2_t*)(dsm + (size_t)s * 2 * CHUNK_B);
      const v2_t* b = (const v2_t*)(dsm + (size_t)s * 2 * CHUNK_B + CHUNK_B);
      v2_t av[2 * VPT], bv[2 * VPT];
#pragma unroll
      for (int j = 0; j < 2 * VPT; ++j) { av[j] = a[j * THREADS + tid]; bv[j] = b[j * THREADS + tid]; }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);  // the stage's data now lives in registers
#pragma unroll
      for (int j = 0; j < 2 * VPT; ++j) { acc0 += expr(av[j].x, bv[j].x); acc1 += expr(av[j].y, bv[j].y); }
    }
  }
  double t = block_sum<THREADS + 32>(acc0 + acc1, smem);
  finish<THREADS + 32>(t, partial, ticket, out, smem, &is_last);
}

struct Variant { const char* name; const void* fn; int threads; void (*launch)(int, const double*, const double*, u64, double*, u32*, double*, cudaStream_t); size_t dsmem; };
template <int THREADS, int MINB, int NACC, bool TRACK, int U>
void launch_v(int grid, const double* A, const double* B, u64 n, double* partial, u32* ticket, double* out, cudaStream_t st) {
  red_kernel<THREADS, MINB, NACC, TRACK, U><<<grid, THREADS, 0, st>>>(A, B, n, partial, ticket, out);
}
#define V(name, T, M, N, K, U) {name, (const void*)red_kernel<T, M, N, K, U>, T, launch_v<T, M, N, K, U>, 0}
template <int THREADS, int MINB>
void launch_p(int grid, const double* A, const double* B, u64 n, double* partial, u32* ticket, double* out, cudaStream_t st) {
  red_prefetch_kernel<THREADS, MINB><<<grid, THREADS, 0, st>>>(A, B, n, partial, ticket, out);
}
#define VP(name, T, M) {name, (const void*)red_prefetch_kernel<T, M>, T, launch_p<T, M>, 0}
template <int THREADS, int MINB, int STAGES, int VPT>
void launch_b(int grid, const double* A, const double* B, u64 n, double* partial, u32* ticket, double* out, cudaStream_t st) {
  red_bulk_kernel<THREADS, MINB, STAGES, VPT><<<grid, THREADS + 32, (size_t)STAGES * 2 * THREADS * VPT * 32, st>>>(A, B, n, partial, ticket, out);
}
#define VB(name, T, M, S, P) {name, (const void*)red_bulk_kernel<T, M, S, P>, T + 32, launch_b<T, M, S, P>, (size_t)S * 2 * T * P * 32}

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
      VP("prefetch   256thr minb4 U1 double-buffered", 256, 4),
      VP("prefetch   256thr minb5 U1 double-buffered", 256, 5),
      VP("prefetch   256thr minb3 U1 double-buffered", 256, 3),
      VB("bulk ring  256thr 4 stages x 16KB  (3/SM)", 256, 3, 4, 1),
      VB("bulk ring  256thr 3 stages x 16KB  (4/SM)", 256, 4, 3, 1),
      VB("bulk ring  256thr 2 stages x 32KB  (3/SM)", 256, 3, 2, 2),
      VB("bulk ring  256thr 3 stages x 32KB  (2/SM)", 256, 2, 3, 2),
      VB("bulk ring  512thr 3 stages x 32KB  (2/SM)", 512, 2, 3, 1),
      VB("bulk ring  512thr 2 stages x 64KB  (1/SM)", 512, 1, 2, 2),
      VB("bulk ring  128thr 4 stages x 8KB   (6/SM)", 128, 6, 4, 1),
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("n = %llu elements (%.1f MB algorithmic per launch), %d SMs, reference sum %.15g\n", n, n * 16.0 / 1e6, sms, ref);
  for (const Variant& v : variants) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, v.fn));
    int per_sm = 0;
    if (v.dsmem) CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.dsmem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, v.threads, v.dsmem));
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
