// gemm_ozaki.cu — matmul engine 2: f64 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) via an
// Ozaki-style integer split.
//
// tcgen05.mma has no f64 kind, but kind::i8 accumulates EXACTLY in int32 inside TMEM. So:
//   1. scale: row i of A by 2^-ea_i, column j of B by 2^-eb_j (exact) so that |x| < 1/2;
//   2. split: x = sum_{s<S} q_s * 128^-(s+1), q_s = rint(.) in [-64,64] (int8), remainder exact in f64 at every step;
//      slices are stored K-contiguous ([S][rows][Kp]) so both operands are K-major for the tensor core;
//   3. multiply: A*B = 2^(ea_i+eb_j) * sum_{d<S} 128^-(d+2) * sum_{s+t=d} A_s * B_t^T.  All pairs of one anti-diagonal d
//      accumulate into ONE int32 TMEM accumulator (|sum| <= S*K*64^2 < 2^31 for K <= 65536), i.e. S(S+1)/2 int8 GEMMs but
//      only S TMEM->register drains per tile;
//   4. recombine: the epilogue warps convert each drained accumulator to f64, scale by the exact power of two and add it
//      into the C tile (smallest anti-diagonal first), applying the row/column exponents (and the MatmulEpilogue) on the
//      last one. The C tile (256 KB) is owned by one CTA, so the read-modify-write stays in L2.
// Truncation error: 2^(-7S) of the row/column maximum per element; S = 7 (49 bits) is below the rounding noise of a
// native f64 dot product of length 8192 (~sqrt(k)*2^-53 relative to sum|a||b|), S = 8 gives 56 bits.
// Round 2: for K <= 21760 the digits are 8 bits wide (base 256, q in [-128, 127] after a carry fix-up, |x| < 0.49): S = 6 slices
// carry 48 bits with 21 + 1 int8 GEMMs instead of 28 + 1 (formulas below with 7 -> 8, 128 -> 256, 2^-16 -> 2^-18).
//
// Element-wise accuracy guard (a posteriori, on the device, no host round trip). The split is accurate relative to
// rowmax(A)_i * colmax(B)_j, not relative to sum_k |a_ik||b_kj| (a row [1e20, 1] against a column [1e-20; 1] loses the
// 1*1 term). In scaled units (|x| < 1/2) the total error of entry (i,j) is at most K * e_S with
//   e_S = 2^-(7S+1)  (digit truncation of both operands)  +  S * 2^-(2+7S)  (dropped anti-diagonals d >= S),
// and  sum_k |xa||xb| >= L_ij * 2^-16  where  L_ij = sum_k |qa0_ik| |qb0_kj|  is ONE more exact int8 GEMM over the absolute
// values of the leading digits (|q| >= 1  =>  128|x| >= |q| - 1/2 >= |q|/2). The kernel computes L_ij as an extra pass
// (slice index S holds |q_0|) and flags every 128x256 tile that contains an entry with  K * e_S > 2^-35 * L_ij * 2^-16;
// flagged tiles are recomputed by the native FP64 (DMMA) kernel, which is launched conditionally on the same stream and
// exits immediately for clean tiles. 2^-35 leaves a factor 3.5 under the 1e-10 * sum|a||b| parity bar. Cost: 29 instead of
// 28 int8 GEMMs. Inf/NaN inputs are detected by the max pass and route every tile to the DMMA kernel the same way.
//
// Kernel structure (persistent, one CTA per SM, 192 threads): warp 0 = TMA producer (4-stage ring of 128x128 A and
// 256x128 B int8 tiles, SWIZZLE_128B), warp 1 = MMA issuer (UMMA 128x256x32, 4 per k-block) + TMEM allocator,
// warps 2-5 = epilogue (tcgen05.ld 32x32b, double-buffered TMEM accumulators so the drain of anti-diagonal d overlaps the
// MMAs of d-1). Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp; developed with scripts/ozaki_dev/i8gemm_test.cu
// (bit-exact vs CPU, 2.41 POP/s at 8192^3).
#include <cuda.h>

#include "common.h"

namespace rm {

namespace {

constexpr int BM = 128, BN = 256, BK = 128, STAGES = 4;
constexpr int A_STAGE = BM * BK, B_STAGE = BN * BK;
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
constexpr int OZ_SMEM = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;  // two 256-column int32 accumulators
// pair kernel: per CTA 128 rows of A + 128 of the pair tile's 256 columns of B per stage
constexpr int P_STAGES = 6;
constexpr int P_B_STAGE = (BN / 2) * BK;
constexpr int P_STAGE_BYTES = A_STAGE + P_B_STAGE;  // 32 KB
constexpr int OZ_SMEM_PAIR = P_STAGES * P_STAGE_BYTES + 1024 + 256;
constexpr int NTHREADS = 192;
constexpr int MAX_SLICES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1); return false; }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO = 8 rows x 128 B
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- cta_group::2 forms (a CTA pair = cluster 2x1x1 owns a 256 x 256 tile; PTX forms as in cute/arch/mma_sm100_umma.hpp,
//      copy_sm100_tma.hpp, cutlass/arch/barrier.h; brought up in scripts/ozaki_dev/i8gemm2_test.cu) --------------------------------
__device__ __forceinline__ void mma_i8_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair when the MMAs issued so far have retired
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// TMA load by either CTA of the pair into ITS OWN shared memory; the bytes are credited to the LEADER's barrier at the same offset
// (CTA-rank bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// arrive (count 1) on the LEADER's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Grouped tile rasterisation: consecutive tile ids walk GROUP_M m-tiles before advancing n, so the ~148 tiles in flight
// form a roughly square patch of the output and share both A row-panels and B row-panels through L2 (r02 ncu: with
// m-fastest order the kernel re-read 32 GB of slices from DRAM per 8192^3 product).
constexpr int GROUP_M = 8;
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int* m0, int* n0) {
  const int per_group = GROUP_M * tiles_n;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  const int r = tile % per_group;
  *m0 = (first_m + r % gsz) * BM;
  *n0 = (r / gsz) * BN;
}

struct OzEpilogue {
  double alpha, beta;
  const double* row_scale;
  const double* col_scale;
  int row_div, col_div, has_min, has_max, has_pow;
  double cmin, cmax, pw;
  double* diag;
  int active;
};
__device__ __forceinline__ double oz_apply_epilogue(double v, const OzEpilogue& ep, uint64_t i, uint64_t j) {
  v = v * ep.alpha + ep.beta;
  if (ep.row_scale) { const double s = ep.row_scale[i]; v = ep.row_div ? v / s : v * s; }
  if (ep.col_scale) { const double s = ep.col_scale[j]; v = ep.col_div ? v / s : v * s; }
  if (ep.has_min) v = fmax(v, ep.cmin);
  if (ep.has_max) v = fmin(v, ep.cmax);
  if (ep.has_pow) v = pow(v, ep.pw);
  if (ep.diag && i == j) ep.diag[i] = v;
  return v;
}

// exponent e with |x| * 2^-e < 1/2 for every |x| <= max (bits = IEEE pattern of max >= 0)
// 8-bit digits additionally need |x| < 0.49, so that the leading digit rint(256 x) stays <= 125 and can absorb a carry (slice_kernel)
__device__ __forceinline__ int scale_exponent(unsigned long long maxbits, int bits) {
  if (maxbits == 0ull) return 0;
  const double mx = __longlong_as_double((long long)maxbits);
  int e = ilogb(mx) + 2;
  if (bits == 8 && scalbn(mx, -e) >= 0.49) ++e;
  return e;
}

// ---- 1. per-row / per-column maxima (as IEEE bit patterns: non-negative doubles order like integers) --------------------
// A is column-major m x k: thread per row, k split over blockIdx.y; B is column-major k x n: warp per column.
__global__ void rowmax_kernel(const double* __restrict__ A, uint64_t m, uint64_t k, unsigned long long* __restrict__ maxbits, int* __restrict__ nonfinite) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t chunk = (k + gridDim.y - 1) / gridDim.y;
  const uint64_t k0 = (uint64_t)blockIdx.y * chunk, k1 = min(k, k0 + chunk);
  double mx = 0.0;
  bool bad = false;
  for (uint64_t kk = k0; kk < k1; ++kk) { const double v = fabs(A[i + kk * m]); bad |= !(v <= 1.7976931348623157e308); mx = fmax(mx, v); }
  if (bad) atomicExch(nonfinite, 1);
  atomicMax(&maxbits[i], (unsigned long long)__double_as_longlong(mx));
}
__global__ void colmax_kernel(const double* __restrict__ B, uint64_t k, uint64_t n, unsigned long long* __restrict__ maxbits, int* __restrict__ nonfinite) {
  const uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  double mx = 0.0;
  bool bad = false;
  for (uint64_t kk = lane; kk < k; kk += 32) { const double v = fabs(B[kk + j * k]); bad |= !(v <= 1.7976931348623157e308); mx = fmax(mx, v); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(nonfinite, 1);
  if (lane == 0) maxbits[j] = (unsigned long long)__double_as_longlong(mx);
}

// ---- 2. slicing: X (column-major) -> int8 slices [S][rows_p][Kp], K contiguous -----------------------------------------
// IS_A: X is m x k, row index = i (contiguous in X) -> transposed on the way out through shared memory.
// !IS_A: X is k x n, row index = j, K already contiguous.
// Tile: 32 rows x 128 k. Each thread converts 16 elements; the write-out is 16-byte vectors, 128 B per (slice,row).
template <bool IS_A>
__global__ void __launch_bounds__(256) slice_kernel(const double* __restrict__ X, uint64_t rows, uint64_t K, const unsigned long long* __restrict__ maxbits,
                                                    int8_t* __restrict__ out, uint64_t rows_p, uint64_t Kp, int S, int bits, const int* __restrict__ flags) {
  extern __shared__ __align__(16) int8_t sh[];  // [S + 1][32][144]; slice S = |q_0| (the accuracy guard's magnitude operand)
  constexpr int PITCH = 144;
  if (flags[0]) return;  // non-finite input: every tile goes to the FP64 kernel, no slices needed
  const uint64_t r0 = (uint64_t)blockIdx.x * 32, k0 = (uint64_t)blockIdx.y * 128;
  const int t = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
    int rl, kl;
    if (IS_A) { rl = t & 31; kl = (t >> 5) + 8 * it; }   // warp = 32 consecutive rows (contiguous in X)
    else { kl = t & 127; rl = (t >> 7) + 2 * it; }        // warp = 32 consecutive k (contiguous in X)
    const uint64_t r = r0 + rl, kk = k0 + kl;
    double x = 0.0;
    if (r < rows && kk < K) {
      const double v = IS_A ? X[r + kk * rows] : X[kk + r * K];
      x = scalbn(v, -scale_exponent(maxbits[r], bits));  // exact; |x| < 1/2 (< 0.49 with 8-bit digits)
    }
    if (bits == 7) {
      for (int s = 0; s < S; ++s) {
        x *= 128.0;                 // exact
        const double q = rint(x);   // |q| <= 64
        x -= q;                     // exact remainder, |x| <= 1/2
        sh[(s * 32 + rl) * PITCH + kl] = (int8_t)(int)q;
        if (s == 0) sh[(S * 32 + rl) * PITCH + kl] = (int8_t)(int)fabs(q);
      }
    } else {
      // base 256: rint gives digits in [-128, 128]; +128 does not fit an int8, so it becomes -128 with a carry into the next
      // more significant digit (128 * 256^-(s+1) = 256^-s - 128 * 256^-(s+1)). The leading digit is <= 125 and absorbs a carry.
      int qd[MAX_SLICES];
#pragma unroll
      for (int s = 0; s < MAX_SLICES; ++s) {
        qd[s] = 0;
        if (s < S) {
          x *= 256.0;                // exact
          const double q = rint(x);  // |q| <= 128
          x -= q;                    // exact remainder, |x| <= 1/2
          qd[s] = (int)q;
        }
      }
#pragma unroll
      for (int s = MAX_SLICES - 1; s >= 1; --s)
        if (qd[s] >= 128) { qd[s] -= 256; qd[s - 1] += 1; }
#pragma unroll
      for (int s = 0; s < MAX_SLICES; ++s)
        if (s < S) sh[(s * 32 + rl) * PITCH + kl] = (int8_t)qd[s];
      sh[(S * 32 + rl) * PITCH + kl] = (int8_t)(qd[0] < 0 ? -qd[0] : qd[0]);
    }
  }
  __syncthreads();
  for (int c = t; c < (S + 1) * 32 * 8; c += 256) {
    const int s = c >> 8, rem = c & 255, rl = rem >> 3, part = rem & 7;
    const uint64_t r = r0 + rl;
    if (r < rows_p)
      *reinterpret_cast<int4*>(out + ((uint64_t)s * rows_p + r) * Kp + k0 + part * 16) = *reinterpret_cast<const int4*>(&sh[(s * 32 + rl) * PITCH + part * 16]);
  }
}

// ---- 3+4. the fused multi-slice tcgen05 GEMM with f64 recombination ------------------------------------------------------
// Passes per tile: pass pi < S is anti-diagonal d = S-1-pi (pairs (s, d-s), s = 0..d), pass S is the accuracy guard's
// magnitude product (the single pair (S, S) of |q_0| slices).
__device__ __forceinline__ int pass_pairs(int pi, int S) { return pi < S ? S - pi : 1; }
__device__ __forceinline__ void pass_pair(int pi, int S, int idx, int* s, int* t) {
  if (pi < S) { *s = idx; *t = (S - 1 - pi) - idx; }
  else { *s = S; *t = S; }
}

__global__ void __launch_bounds__(NTHREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C, uint64_t M, uint64_t N,
                  int Mp, int Np, int Kp, int S, int bits, const unsigned long long* __restrict__ amax, const unsigned long long* __restrict__ bmax,
                  const __grid_constant__ OzEpilogue ep, int* __restrict__ flags, int* __restrict__ tileflags, long long guard_min) {
  extern __shared__ uint8_t smem_raw[];
  if (flags[0]) return;  // non-finite input (uniform for the grid): the conditional FP64 kernel computes every tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  int* err = flags + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = Mp / BM, tiles_n = Np / BN, ntiles = tiles_m * tiles_n;
  const int nk = Kp / BK;
  const int npass = S + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter -> stage / phase
      bool ok = true;
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
        int m0, n0;
        tile_coords(tile, tiles_m, tiles_n, &m0, &n0);
        for (int pi = 0; pi < npass && ok; ++pi) {
          const int np = pass_pairs(pi, S);
          for (int idx = 0; idx < np && ok; ++idx) {
            int s, t;
            pass_pair(pi, S, idx, &s, &t);
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int st = it % STAGES;
              if (!mbar_wait(&empty[st], ((it / STAGES) & 1) ^ 1, err)) { ok = false; break; }
              mbar_expect_tx(&full[st], STAGE_BYTES);
              tma_load_2d(smem + st * STAGE_BYTES, &tmA, &full[st], kb * BK, s * Mp + m0);
              tma_load_2d(smem + st * STAGE_BYTES + A_STAGE, &tmB, &full[st], kb * BK, t * Np + n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, g = 0;  // g: global pass counter -> TMEM buffer / phase
      bool ok = true;
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
        for (int pi = 0; pi < npass && ok; ++pi, ++g) {
          const uint32_t buf = g & 1;
          if (!mbar_wait(&tmem_empty[buf], ((g >> 1) & 1) ^ 1, err)) { ok = false; break; }
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          uint32_t first = 1;
          const int np = pass_pairs(pi, S);
          for (int idx = 0; idx < np && ok; ++idx)
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int st = it % STAGES;
              if (!mbar_wait(&full[st], (it / STAGES) & 1, err)) { ok = false; break; }
              tc_fence_after();
              const uint32_t a_addr = smem_u32(smem + st * STAGE_BYTES), b_addr = a_addr + A_STAGE;
              const uint64_t da = make_desc_k_sw128(a_addr), db = make_desc_k_sw128(b_addr);
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) { mma_i8(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u); first = 0; }
              tc_commit(&empty[st]);
            }
          if (ok) tc_commit(&tmem_full[buf]);
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t g = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
      int m0, n0;
      tile_coords(tile, tiles_m, tiles_n, &m0, &n0);
      const uint64_t row = (uint64_t)m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const int ea = row_ok ? scale_exponent(amax[row], bits) : 0;
      for (int pi = 0; pi < npass && ok; ++pi, ++g) {
        const uint32_t buf = g & 1;
        if (!mbar_wait(&tmem_full[buf], (g >> 1) & 1, err)) { ok = false; break; }
        tc_fence_after();
        if (pi < S) {
          const int d = S - 1 - pi;
          const double scale = scalbn(1.0, -bits * (d + 2));
          const bool first = pi == 0, last = d == 0;
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c, v);
            if (row_ok) {
              double acc[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint64_t col = (uint64_t)n0 + c + j;
                acc[j] = (!first && col < N) ? C[row + col * M] : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint64_t col = (uint64_t)n0 + c + j;
                if (col < N) {
                  double r = acc[j] + (double)(int)v[j] * scale;
                  if (last) {
                    r = scalbn(r, ea + scale_exponent(bmax[col], bits));
                    if (ep.active) r = oz_apply_epilogue(r, ep, row, col);
                  }
                  C[row + col * M] = r;
                }
              }
            }
          }
        } else {
          // accuracy guard: L_ij = sum_k |qa0||qb0| must reach guard_min, else the tile is recomputed in native f64
          bool weak = false;
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) weak |= row_ok && ((uint64_t)n0 + c + j < N) && ((long long)(int)v[j] < guard_min);
          }
          if (__any_sync(0xffffffffu, weak) && lane == 0) {
            if (atomicExch(&tileflags[m0 / BM + (n0 / BN) * tiles_m], 1) == 0) atomicAdd(&flags[2], 1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// The same pipeline on CTA PAIRS (cta_group::2, cluster 2x1x1): a pair owns a 256 x 256 tile, each CTA stages its 128 rows of the A
// slice and its 128-column half of the B slice (32 KB per stage instead of 48 KB, 6 stages), the leader's MMA warp issues UMMA
// 256x256x32 for both, TMA completions of both CTAs are credited to the leader's `full` barrier, stage releases and accumulator-ready
// signals are multicast commits, the epilogue warps of both CTAs drain their own 128 TMEM lanes and release the accumulator on the
// leader's `tmem_empty` barrier (count 8). Shared-memory fill traffic per SM and MMA drops by a third, which is what limited the
// single-CTA mainloop (r27 harness, one 8192^3 int8 GEMM: 2.53 vs 2.31 POP/s). Mp is padded to a multiple of 256.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
ozaki_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C, uint64_t M, uint64_t N,
                  int Mp, int Np, int Kp, int S, int bits, const unsigned long long* __restrict__ amax, const unsigned long long* __restrict__ bmax,
                  const __grid_constant__ OzEpilogue ep, int* __restrict__ flags, int* __restrict__ tileflags, long long guard_min) {
  extern __shared__ uint8_t smem_raw[];
  if (flags[0]) return;  // non-finite input (uniform for the grid): the conditional FP64 kernel computes every tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* empty = full + P_STAGES;
  uint64_t* tmem_full = empty + P_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  int* err = flags + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader of the pair
  const int tiles_m = Mp / (2 * BM), tiles_n = Np / BN, ntiles = tiles_m * tiles_n;  // pair tiles: 256 x 256
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk = Kp / BK;
  const int npass = S + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }  // 4 epilogue warps x 2 CTAs release an accumulator
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before a remote completion or a multicast commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter -> stage / phase
      bool ok = true;
      for (int tile = pair_id; tile < ntiles && ok; tile += npairs) {
        int m0, n0;
        tile_coords(tile, tiles_m, tiles_n, &m0, &n0);
        m0 = 2 * m0 + (int)rank * BM;  // tile_coords counts in BM rows; pair tiles are 2 BM tall
        for (int pi = 0; pi < npass && ok; ++pi) {
          const int np = pass_pairs(pi, S);
          for (int idx = 0; idx < np && ok; ++idx) {
            int s, t;
            pass_pair(pi, S, idx, &s, &t);
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int st = it % P_STAGES;
              if (!mbar_wait(&empty[st], ((it / P_STAGES) & 1) ^ 1, err)) { ok = false; break; }
              if (rank == 0) mbar_expect_tx(&full[st], 2 * P_STAGE_BYTES);  // the leader's barrier collects the bytes of both CTAs
              tma_load_2d_pair(smem + st * P_STAGE_BYTES, &tmA, &full[st], kb * BK, s * Mp + m0);
              tma_load_2d_pair(smem + st * P_STAGE_BYTES + A_STAGE, &tmB, &full[st], kb * BK, t * Np + n0 + (int)rank * (BN / 2));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);  // M = 256
      uint32_t it = 0, g = 0;  // g: global pass counter -> TMEM buffer / phase
      bool ok = true;
      for (int tile = pair_id; tile < ntiles && ok; tile += npairs) {
        for (int pi = 0; pi < npass && ok; ++pi, ++g) {
          const uint32_t buf = g & 1;
          if (!mbar_wait(&tmem_empty[buf], ((g >> 1) & 1) ^ 1, err)) { ok = false; break; }
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          uint32_t first = 1;
          const int np = pass_pairs(pi, S);
          for (int idx = 0; idx < np && ok; ++idx)
            for (int kb = 0; kb < nk; ++kb, ++it) {
              const int st = it % P_STAGES;
              if (!mbar_wait(&full[st], (it / P_STAGES) & 1, err)) { ok = false; break; }
              tc_fence_after();
              const uint32_t a_addr = smem_u32(smem + st * P_STAGE_BYTES), b_addr = a_addr + A_STAGE;
              const uint64_t da = make_desc_k_sw128(a_addr), db = make_desc_k_sw128(b_addr);
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) { mma_i8_pair(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u); first = 0; }
              tc_commit_pair(&empty[st]);
            }
          if (ok) tc_commit_pair(&tmem_full[buf]);
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    uint32_t g = 0;
    bool ok = true;
    for (int tile = pair_id; tile < ntiles && ok; tile += npairs) {
      int m0, n0;
      tile_coords(tile, tiles_m, tiles_n, &m0, &n0);
      m0 = 2 * m0 + (int)rank * BM;
      const uint64_t row = (uint64_t)m0 + q * 32 + lane;
      const bool row_ok = row < M;
      const int ea = row_ok ? scale_exponent(amax[row], bits) : 0;
      for (int pi = 0; pi < npass && ok; ++pi, ++g) {
        const uint32_t buf = g & 1;
        if (!mbar_wait(&tmem_full[buf], (g >> 1) & 1, err)) { ok = false; break; }
        tc_fence_after();
        if (pi < S) {
          const int d = S - 1 - pi;
          const double scale = scalbn(1.0, -bits * (d + 2));
          const bool first = pi == 0, last = d == 0;
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c, v);
            if (row_ok) {
              double acc[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint64_t col = (uint64_t)n0 + c + j;
                acc[j] = (!first && col < N) ? C[row + col * M] : 0.0;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const uint64_t col = (uint64_t)n0 + c + j;
                if (col < N) {
                  double r = acc[j] + (double)(int)v[j] * scale;
                  if (last) {
                    r = scalbn(r, ea + scale_exponent(bmax[col], bits));
                    if (ep.active) r = oz_apply_epilogue(r, ep, row, col);
                  }
                  C[row + col * M] = r;
                }
              }
            }
          }
        } else {
          // accuracy guard: L_ij = sum_k |qa0||qb0| must reach guard_min, else the tile is recomputed in native f64
          bool weak = false;
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) weak |= row_ok && ((uint64_t)n0 + c + j < N) && ((long long)(int)v[j] < guard_min);
          }
          if (__any_sync(0xffffffffu, weak) && lane == 0) {
            if (atomicExch(&tileflags[m0 / BM + (n0 / BN) * (Mp / BM)], 1) == 0) atomicAdd(&flags[2], 1);  // flags stay per 128 x 256 tile
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory and TMEM stay alive until the leader's last MMA has retired and both epilogues are done
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    cudaGetLastError();
  });
  return fn;
}

rm_status make_map(void* ptr, uint64_t rows, uint64_t k, uint32_t box_rows, CUtensorMap* out) {
  EncodeTiledFn enc = encode_tiled();
  RM_REQUIRE(enc != nullptr, RM_UNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {k};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RM_REQUIRE(r == CUDA_SUCCESS, RM_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RM_OK;
}

}  // namespace

int ozaki_default_slices() {
  if (const char* e = getenv("RUNMAT_B200_OZAKI_SLICES")) {
    int s = atoi(e);
    if (s >= 2 && s <= MAX_SLICES) return s;
  }
  return 7;
}

// Per-provider workspace: slice buffers, exponent arrays, device flags and the two tensor maps are kept between calls (an
// 8192^3 product needs ~1 GB of slices; re-allocating them and re-encoding the maps per call was pure host overhead).
// Guarded by rm_provider::oz_mu, which matmul_impl holds from the first enqueue of a product to its conditional FP64 launch.
struct OzWorkspace {
  int8_t *As = nullptr, *Bs = nullptr;
  size_t as_bytes = 0, bs_bytes = 0;
  unsigned long long *amax = nullptr, *bmax = nullptr;
  size_t amax_n = 0, bmax_n = 0;
  int* flags = nullptr;      // [0] non-finite input, [1] kernel protocol error, [2] tiles handed to the FP64 kernel, [3] spare
  int* tileflags = nullptr;  // [tiles_m * tiles_n], m fastest
  size_t tile_cap = 0;
  bool attrs_set = false;
  int attr_slices = 0;
  CUtensorMap tmA, tmB;
  void *mapA_ptr = nullptr, *mapB_ptr = nullptr;
  uint64_t mapA_rows = 0, mapB_rows = 0, mapA_k = 0, mapB_k = 0;
  uint32_t mapB_box = 0;
  int last_gemms = 0;  // int8 GEMMs per output tile of the last product: S(S+1)/2 digit products + the guard's magnitude product
};

void ozaki_workspace_destroy(rm_provider* p) {
  OzWorkspace* w = (OzWorkspace*)p->oz_ws;
  if (!w) return;
  for (void* q : {(void*)w->As, (void*)w->Bs, (void*)w->amax, (void*)w->bmax, (void*)w->flags, (void*)w->tileflags})
    if (q) cudaFreeAsync(q, p->stream);
  delete w;
  p->oz_ws = nullptr;
}

namespace {
template <typename T>
cudaError_t grow(rm_provider* p, T** ptr, size_t* have, size_t want) {
  if (want <= *have) return cudaSuccess;
  if (*ptr) { cudaError_t e = cudaFreeAsync(*ptr, p->stream); if (e != cudaSuccess) return e; }
  *ptr = nullptr;
  *have = 0;
  cudaError_t e = cudaMallocAsync((void**)ptr, want * sizeof(T), p->stream);
  if (e == cudaSuccess) *have = want;
  return e;
}
}  // namespace

// Enqueues the Ozaki/tcgen05 product of A (m x k) and B (k x n) into C. No host synchronisation: non-finite inputs and
// tiles that fail the accuracy guard are marked in device flags, and the caller (matmul_impl) enqueues the FP64 kernel
// conditionally on `guard`. *used = false (RM_OK) only when the shape is outside the engine's range.
// The caller must hold p->oz_mu until its conditional launch is enqueued.
rm_status ozaki_matmul(rm_provider* p, const double* A, const double* B, double* C, uint64_t m, uint64_t n, uint64_t k,
                       const rm_matmul_epilogue* epd, const void* prow, const void* pcol, void* pdiag, bool ep_active, bool* used, OzGuard* guard) {
  *used = false;
  if (k > 65536 || m == 0 || n == 0 || k == 0) return RM_OK;  // int32 headroom of the 7-bit split: S*K*64^2 < 2^31
  // CTA pairs (cta_group::2, 256 x 256 tiles) unless switched off; Mp is then a multiple of 256
  const bool pair = getenv("RUNMAT_B200_OZAKI_1CTA") == nullptr;
  const uint64_t MT = pair ? 2 * BM : BM;
  const uint64_t Mp = (m + MT - 1) / MT * MT, Np = (n + BN - 1) / BN * BN, Kp = (k + BK - 1) / BK * BK;
  // Digit width. 8-bit digits (|q| <= 128) cover 48 bits with S = 6 slices = 21 int8 GEMMs instead of the 28 of the 7-bit split
  // (S = 7, 49 bits); an anti-diagonal of up to 6 pairs then needs 6 * K * 128^2 < 2^31, i.e. K <= 21760. Longer products keep
  // the 7-bit digits. RUNMAT_B200_OZAKI_BITS / RUNMAT_B200_OZAKI_SLICES override.
  int bits = Kp * 6ull * 16384ull < (1ull << 31) ? 8 : 7;
  if (const char* e = getenv("RUNMAT_B200_OZAKI_BITS")) { const int b = atoi(e); if (b == 7 || (b == 8 && bits == 8)) bits = b; }
  int S = bits == 8 ? 6 : 7;
  if (getenv("RUNMAT_B200_OZAKI_SLICES")) S = ozaki_default_slices();
  if (bits == 8 && (uint64_t)S * Kp * 16384ull >= (1ull << 31)) { bits = 7; }
  if (Mp * (S + 1) >= (1ull << 31) || Np * (S + 1) >= (1ull << 31)) return RM_OK;
  cudaStream_t st = p->stream;
  if (!p->oz_ws) p->oz_ws = new OzWorkspace();
  OzWorkspace& w = *(OzWorkspace*)p->oz_ws;
#define OZ_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); return fail(_e == cudaErrorMemoryAllocation ? RM_OOM : RM_ERROR, "%s failed: %s", #expr, cudaGetErrorString(_e)); } } while (0)
  const size_t ntiles = (size_t)(Mp / BM) * (Np / BN);
  OZ_CUDA(grow(p, &w.amax, &w.amax_n, (size_t)Mp));
  OZ_CUDA(grow(p, &w.bmax, &w.bmax_n, (size_t)Np));
  OZ_CUDA(grow(p, &w.tileflags, &w.tile_cap, ntiles));
  if (!w.flags) OZ_CUDA(cudaMallocAsync((void**)&w.flags, 4 * sizeof(int), st));
  OZ_CUDA(grow(p, &w.As, &w.as_bytes, (size_t)(S + 1) * Mp * Kp));
  OZ_CUDA(grow(p, &w.Bs, &w.bs_bytes, (size_t)(S + 1) * Np * Kp));
  OZ_CUDA(cudaMemsetAsync(w.amax, 0, Mp * 8, st));
  OZ_CUDA(cudaMemsetAsync(w.flags, 0, 4 * sizeof(int), st));
  OZ_CUDA(cudaMemsetAsync(w.tileflags, 0, ntiles * sizeof(int), st));
  {
    const unsigned ky = (unsigned)std::min<uint64_t>(64, std::max<uint64_t>(1, k / 256));
    rowmax_kernel<<<dim3((unsigned)((m + 255) / 256), ky), 256, 0, st>>>(A, m, k, w.amax, w.flags);
    colmax_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(B, k, n, w.bmax, w.flags);
  }
  const size_t slice_smem = (size_t)(S + 1) * 32 * 144;
  if (!w.attrs_set || w.attr_slices != S) {
    OZ_CUDA(cudaFuncSetAttribute(slice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_smem));
    OZ_CUDA(cudaFuncSetAttribute(slice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_smem));
    OZ_CUDA(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
    OZ_CUDA(cudaFuncSetAttribute(ozaki_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_PAIR));
    w.attrs_set = true;
    w.attr_slices = S;
  }
  slice_kernel<true><<<dim3((unsigned)(Mp / 32), (unsigned)(Kp / 128)), 256, slice_smem, st>>>(A, m, k, w.amax, w.As, Mp, Kp, S, bits, w.flags);
  slice_kernel<false><<<dim3((unsigned)(Np / 32), (unsigned)(Kp / 128)), 256, slice_smem, st>>>(B, n, k, w.bmax, w.Bs, Np, Kp, S, bits, w.flags);
  if (w.mapA_ptr != w.As || w.mapA_rows != (uint64_t)(S + 1) * Mp || w.mapA_k != Kp) {
    RM_TRY(make_map(w.As, (uint64_t)(S + 1) * Mp, Kp, BM, &w.tmA));
    w.mapA_ptr = w.As; w.mapA_rows = (uint64_t)(S + 1) * Mp; w.mapA_k = Kp;
  }
  const uint32_t boxB = pair ? BN / 2 : BN;  // a CTA of a pair stages its 128-column half of the B tile
  if (w.mapB_ptr != w.Bs || w.mapB_rows != (uint64_t)(S + 1) * Np || w.mapB_k != Kp || w.mapB_box != boxB) {
    RM_TRY(make_map(w.Bs, (uint64_t)(S + 1) * Np, Kp, boxB, &w.tmB));
    w.mapB_ptr = w.Bs; w.mapB_rows = (uint64_t)(S + 1) * Np; w.mapB_k = Kp; w.mapB_box = boxB;
  }
  OzEpilogue ep{};
  ep.alpha = 1.0;
  if (epd)
    ep = OzEpilogue{epd->alpha, epd->beta, (const double*)prow, (const double*)pcol, epd->row_op == RM_SCALE_DIVIDE, epd->col_op == RM_SCALE_DIVIDE,
                    epd->has_clamp_min, epd->has_clamp_max, epd->has_pow, epd->clamp_min, epd->clamp_max, epd->pow_exponent, (double*)pdiag, ep_active ? 1 : 0};
  // accuracy guard threshold (see the header comment), digit base B = 2^bits: flag when K * e_S > 2^-35 * L_ij * (2B)^-2 with
  // e_S = B^-S * (1/2 + S/4), i.e. L_ij must reach K * (1/2 + S/4) * 2^(37 + 2 bits - bits S)   (bits = 7: 2^(51 - 7S))
  const double gmin = std::ceil((double)k * std::ldexp(0.5 + 0.25 * S, 37 + 2 * bits - bits * S));
  const long long guard_min = gmin >= 4.0e18 ? (long long)4e18 : std::max<long long>(1, (long long)gmin);
  if (pair) {
    const size_t npair_tiles = (size_t)(Mp / (2 * BM)) * (Np / BN);
    const int grid = 2 * (int)std::min<size_t>(npair_tiles, (size_t)p->prop.multiProcessorCount / 2);
    ozaki_gemm_pair_kernel<<<grid, NTHREADS, OZ_SMEM_PAIR, st>>>(w.tmA, w.tmB, C, m, n, (int)Mp, (int)Np, (int)Kp, S, bits, w.amax, w.bmax, ep, w.flags, w.tileflags, guard_min);
  } else {
    const int grid = (int)std::min<size_t>(ntiles, (size_t)p->prop.multiProcessorCount);
    ozaki_gemm_kernel<<<grid, NTHREADS, OZ_SMEM, st>>>(w.tmA, w.tmB, C, m, n, (int)Mp, (int)Np, (int)Kp, S, bits, w.amax, w.bmax, ep, w.flags, w.tileflags, guard_min);
  }
  OZ_CUDA(cudaGetLastError());
  count_launch(p, 5);
#undef OZ_CUDA
  w.last_gemms = S * (S + 1) / 2 + 1;
  guard->flags = w.flags;
  guard->tileflags = w.tileflags;
  guard->tiles_m = (int)(Mp / BM);
  *used = true;
  return RM_OK;
}

// Test / debug hook: waits for the stream and returns the device flags of the LAST Ozaki product
// (out[0] non-finite input, out[1] pipeline protocol error, out[2] tiles recomputed by the FP64 kernel).
rm_status ozaki_last_stats(rm_provider* p, int out[4]) {
  std::lock_guard<std::mutex> lk(p->oz_mu);
  OzWorkspace* w = (OzWorkspace*)p->oz_ws;
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!w || !w->flags) return RM_OK;
  RM_CUDA(cudaMemcpyAsync(out, w->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  RM_CUDA(cudaStreamSynchronize(p->stream));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  out[3] = w->last_gemms;
  return RM_OK;
}

}  // namespace rm
