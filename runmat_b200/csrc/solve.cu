// solve.cu — mldivide (row a9): X = A \ B for square A on the device.
//
// Reference: the host solves through an SVD (nalgebra 0.32.6, builtins/math/linalg/ops/mldivide.rs:317-397) and the wgpu
// provider simply downloads, solves on the host and re-uploads (backend/wgpu/provider/ops/solve.rs:131-205). The
// reference's own tests pin the result by residual only (mldivide.rs:662-676: ||A*X - B|| < 1e-12), so an LU solve is
// within contract for well-conditioned square systems. Everything else (non-square / singular / ill-conditioned) returns
// RM_UNSUPPORTED so the caller falls back to the host SVD path exactly as it does today.
//
// B200 design: right-looking blocked LU with partial pivoting, NB = 64.
//   panel   one cooperative kernel per panel (grid.sync between the pivot search and the rank-1 update of each column);
//           the pivot search of column c+1 is fused into the update pass of column c;
//   laswp   one thread per column applies the panel's NB row interchanges to the rest of LU and to the right-hand side;
//   trsm    64 x 64 triangular blocks solved in shared memory, 32 right-hand-side columns per CTA;
//   update  A22 -= A21 * A12 and the block forward/backward substitutions run on the FP64 tensor-core GEMM
//           (dgemm_dmma_kernel with leading dimensions, gemm.cu).
#include <cooperative_groups.h>

#include "common.h"
#include "lu_panel_push.h"

namespace cg = cooperative_groups;

namespace rm {

namespace {

constexpr int NB = 64;

struct PivotEntry {
  double val;
  unsigned long long idx;
};

// Factor the panel A[j0:n, j0:j0+jb] (column-major, lda) in place. ipiv[j0+c] = global row swapped with row j0+c.
// info: set to 1 when a pivot is exactly zero (singular), min_piv/max_abs feed the conditioning check on the host.
__global__ void __launch_bounds__(256)
lu_panel_kernel(double* __restrict__ A, uint64_t lda, uint64_t n, uint64_t j0, int jb, unsigned long long* __restrict__ ipiv,
                PivotEntry* __restrict__ scratch, int* __restrict__ info, double* __restrict__ piv_minmax) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_val[256];
  __shared__ unsigned long long s_idx[256];
  __shared__ double s_row[NB];
  __shared__ unsigned long long s_piv;
  const uint64_t m = n - j0;  // panel rows
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
  double* P = A + j0 + j0 * lda;  // panel origin

  // local candidate for column 0
  double best = -1.0;
  unsigned long long best_i = 0;
  for (uint64_t r = tid; r < m; r += nthr) { const double v = fabs(P[r]); if (v > best) { best = v; best_i = r; } }

  for (int c = 0; c < jb; ++c) {
    // ---- block-level argmax (ties -> smallest row: deterministic) ----
    s_val[threadIdx.x] = best;
    s_idx[threadIdx.x] = best_i;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) {
        const double ov = s_val[threadIdx.x + off];
        const unsigned long long oi = s_idx[threadIdx.x + off];
        if (ov > s_val[threadIdx.x] || (ov == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) { s_val[threadIdx.x] = ov; s_idx[threadIdx.x] = oi; }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) { scratch[blockIdx.x].val = s_val[0]; scratch[blockIdx.x].idx = s_idx[0]; }
    grid.sync();
    // ---- every block resolves the global pivot redundantly; block 0 swaps the two rows inside the panel ----
    if (threadIdx.x == 0) {
      double bv = -1.0;
      unsigned long long bi = 0;
      for (unsigned b = 0; b < gridDim.x; ++b) {
        const double v = scratch[b].val;
        const unsigned long long i = scratch[b].idx;
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
      s_piv = bi;
      if (blockIdx.x == 0) {
        ipiv[j0 + c] = j0 + bi;
        if (!(bv > 0.0)) atomicExch(info, 1);
        // track min/max |pivot| for the conditioning estimate
        if (bv > 0.0) {
          if (bv < piv_minmax[0]) piv_minmax[0] = bv;
          if (bv > piv_minmax[1]) piv_minmax[1] = bv;
        }
      }
    }
    __syncthreads();
    const uint64_t prow = s_piv;
    if (blockIdx.x == 0 && prow != (uint64_t)c) {
      for (int cc = threadIdx.x; cc < jb; cc += blockDim.x) {
        const double a = P[c + (uint64_t)cc * lda], b = P[prow + (uint64_t)cc * lda];
        P[c + (uint64_t)cc * lda] = b;
        P[prow + (uint64_t)cc * lda] = a;
      }
    }
    grid.sync();
    // ---- scale column c and rank-1 update of the remaining panel columns; fused pivot search for column c+1 ----
    for (int cc = threadIdx.x; cc < jb; cc += blockDim.x) s_row[cc] = P[c + (uint64_t)cc * lda];
    __syncthreads();
    const double pivot = s_row[c];
    best = -1.0;
    best_i = 0;
    if (pivot != 0.0) {
      for (uint64_t r = (uint64_t)c + 1 + tid; r < m; r += nthr) {
        const double l = P[r + (uint64_t)c * lda] / pivot;
        P[r + (uint64_t)c * lda] = l;
        int cc = c + 1;
        if (cc < jb) {  // next column: also the pivot candidate for step c+1
          const double v = P[r + (uint64_t)cc * lda] - l * s_row[cc];
          P[r + (uint64_t)cc * lda] = v;
          const double av = fabs(v);
          if (av > best) { best = av; best_i = r; }
          ++cc;
        }
        for (; cc + 7 < jb; cc += 8) {  // 8 independent loads in flight before the stores (the loop is L2-latency bound)
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = P[r + (uint64_t)(cc + u) * lda];
#pragma unroll
          for (int u = 0; u < 8; ++u) P[r + (uint64_t)(cc + u) * lda] = v[u] - l * s_row[cc + u];
        }
        for (; cc < jb; ++cc) P[r + (uint64_t)cc * lda] = P[r + (uint64_t)cc * lda] - l * s_row[cc];
      }
    } else {
      for (uint64_t r = (uint64_t)c + 1 + tid; r < m; r += nthr)
        if (c + 1 < jb) { const double av = fabs(P[r + (uint64_t)(c + 1) * lda]); if (av > best) { best = av; best_i = r; } }
    }
    __syncthreads();
  }
}

// Shared-memory variant (used whenever the panel's rows fit 256 per resident CTA): each CTA keeps its 256 x NB slab of the
// panel in shared memory (one row per thread, padded: conflict-free), so the rank-1 updates never leave the SM. Only the
// pivot search result and the two exchanged rows cross CTAs (global scratch + grid.sync): 2 grid syncs per column and
// no dependent L2 round trips in the update (the first version spent ~6 us per column waiting on them).
constexpr int SLAB_ROWS = 256;
// RowMoves (net row permutation of a panel) lives in lu_panel_push.h, shared with the register-resident cluster kernel.
static_assert(NB == LU_NB, "panel width");
// Per-CTA candidate published before the (single) grid.sync of a column: local max |a(r,c)|, its row, that row's panel
// values, and — from the CTA that owns row c — row c itself (it moves to the pivot's position).
struct Candidate {
  double val;
  unsigned long long idx;
  double row[NB];
};
template <bool CLUSTER>
__global__ void __launch_bounds__(SLAB_ROWS)
lu_panel_smem_kernel(double* __restrict__ A, uint64_t lda, uint64_t n, uint64_t j0, int jb, unsigned long long* __restrict__ ipiv,
                     Candidate* __restrict__ cand /*[2][gridDim.x]*/, double* __restrict__ rowc /*[2][NB]*/, int* __restrict__ info,
                     double* __restrict__ piv_minmax, RowMoves* __restrict__ moves) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double slab_smem[];
  double (*slab)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(slab_smem);  // [SLAB_ROWS][NB+1]
  __shared__ double s_wval[SLAB_ROWS / 32];
  __shared__ unsigned long long s_widx[SLAB_ROWS / 32];
  __shared__ double s_row[NB], s_rowc[NB];
  __shared__ unsigned long long s_piv;
  __shared__ unsigned s_pblock;
  // net row permutation of this panel, maintained by warp 0 of CTA 0 (rows touched <= 2*NB)
  __shared__ unsigned long long p_rows[2 * NB], p_cur[2 * NB];
  __shared__ unsigned p_n;
  __shared__ unsigned long long s_pivots[NB];
  // CLUSTER variant: the per-column exchange (candidates, row c) stays in distributed shared memory and the per-column
  // barrier is a cluster barrier (~0.3 us) instead of a grid-wide cooperative sync (~2.5 us). One cluster = the whole grid.
  __shared__ Candidate s_cand[2];
  __shared__ double s_pub_rowc[2][NB];
  // CLUSTER variant, per-column exchange without the hardware cluster barrier (r17 ncu: barrier.cluster wait = 29 % of the kernel's
  // stall samples, the remote candidate reads that followed it another 13 %): every CTA PUSHES its candidate (|value|, row) into
  // slot [buf][rank] of every peer's mailbox with plain DSMEM stores followed by a release store of the column stamp; a CTA then
  // only spins (acquire) on its OWN shared memory until all stamps of the column have arrived and resolves the pivot locally.
  // Double buffering is safe: a CTA reaches column c+2 only after it has seen every peer's stamp for c+1, which a peer sends after
  // it has finished reading column c.
  __shared__ double mb_val[2][16];
  __shared__ unsigned long long mb_idx[2][16];
  __shared__ unsigned mb_flag[2][16];
  const unsigned nblk = gridDim.x;
  const uint64_t m = n - j0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t r = (uint64_t)blockIdx.x * SLAB_ROWS + tid;  // panel-local row owned by this thread
  const bool valid = r < m;
  double* P = A + j0 + j0 * lda;
  for (int cc = 0; cc < jb; ++cc) slab[tid][cc] = valid ? P[r + (uint64_t)cc * lda] : 0.0;
  if (blockIdx.x == 0 && tid == 0) p_n = 0;
  if constexpr (CLUSTER) {
    if (tid < 32) { mb_flag[0][tid & 15] = 0; mb_flag[1][tid & 15] = 0; }
    cg::this_cluster().sync();  // every mailbox is initialised before the first push
  }
  __syncthreads();

  double loc_pmin = 1.7976931348623157e308, loc_pmax = 0.0;  // CTA 0 / lane 0: extreme |pivot| of this panel (a global read-modify-write per column sat on the critical path)
  for (int c = 0; c < jb; ++c) {
    const int buf = c & 1;  // double-buffered exchange area: column c+1 may be published while a slow CTA still reads column c
    Candidate* mine = CLUSTER ? &s_cand[buf] : cand + (size_t)buf * nblk + blockIdx.x;
    // ---- local argmax over this CTA's rows >= c: warp shuffles, then one shared hop (ties -> smallest row) ----
    double bv = (valid && r >= (uint64_t)c) ? fabs(slab[tid][c]) : -1.0;
    unsigned long long bi = r;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const unsigned long long oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_wval[warp] = bv; s_widx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < SLAB_ROWS / 32 ? s_wval[lane] : -1.0;
      bi = lane < SLAB_ROWS / 32 ? s_widx[lane] : ~0ull;
#pragma unroll
      for (int off = SLAB_ROWS / 64; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const unsigned long long oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { s_piv = bi; mine->val = bv; mine->idx = bi; }
    }
    __syncthreads();
    // publish the candidate row and (if owned here) row c BEFORE the sync, so one barrier per column is enough. One thread
    // per panel column: a single thread walking the 64-wide row cost ~1 us per column of the factorisation.
    {
      const uint64_t lp = s_piv - (uint64_t)blockIdx.x * SLAB_ROWS;  // local row of this CTA's candidate
      if (tid < jb && lp < (uint64_t)SLAB_ROWS) mine->row[tid] = slab[lp][tid];
      if (blockIdx.x == 0 && tid >= NB && tid < NB + jb) (CLUSTER ? s_pub_rowc[buf] : rowc + buf * NB)[tid - NB] = slab[c][tid - NB];
    }
    if constexpr (CLUSTER) {
      __syncthreads();  // this CTA's candidate row (and row c) are complete in its shared memory before the stamp goes out
      if (warp == 0 && (unsigned)lane < nblk) {
        cg::cluster_group cluster = cg::this_cluster();
        const unsigned me = blockIdx.x;
        const double myv = s_cand[buf].val;
        const unsigned long long myi = s_cand[buf].idx;
        *cluster.map_shared_rank(&mb_val[buf][me], lane) = myv;
        *cluster.map_shared_rank(&mb_idx[buf][me], lane) = myi;
        unsigned* fl = cluster.map_shared_rank(&mb_flag[buf][me], lane);
        asm volatile("st.release.cluster.u32 [%0], %1;" ::"l"(fl), "r"((unsigned)(c + 1)) : "memory");
        // wait for peer `lane`'s stamp in OUR mailbox (bounded: a protocol bug raises info instead of hanging the GPU)
        const unsigned* mine_fl = &mb_flag[buf][lane];
        const long long t0 = clock64();
        for (;;) {
          unsigned f;
          asm volatile("ld.acquire.cluster.u32 %0, [%1];" : "=r"(f) : "l"(mine_fl) : "memory");
          if (f == (unsigned)(c + 1)) break;
          if (clock64() - t0 > 4000000000LL) { atomicExch(info, 1); break; }
        }
      }
      __syncwarp();
    } else {
      grid.sync();
    }
    // ---- every CTA resolves the global pivot redundantly: warp 0, one candidate per lane ----
    if (warp == 0) {
      double gv = -1.0;
      unsigned long long gi = ~0ull;
      unsigned gb = 0;
      for (unsigned b = lane; b < nblk; b += 32) {
        double v;
        unsigned long long i;
        if constexpr (CLUSTER) {
          v = *(volatile double*)&mb_val[buf][b]; i = *(volatile unsigned long long*)&mb_idx[buf][b];
        } else {
          const Candidate* cb = cand + (size_t)buf * nblk + b;
          v = __ldcg(&cb->val); i = __ldcg(&cb->idx);
        }
        if (v > gv || (v == gv && i < gi)) { gv = v; gi = i; gb = b; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, gv, off);
        const unsigned long long oi = __shfl_xor_sync(0xffffffffu, gi, off);
        const unsigned ob = __shfl_xor_sync(0xffffffffu, gb, off);
        if (ov > gv || (ov == gv && oi < gi)) { gv = ov; gi = oi; gb = ob; }
      }
      if (gi == ~0ull) gi = (unsigned long long)c;  // NaN column: no candidate compared greater; keep the swap in bounds (the solve is rejected later)
      if (lane == 0) { s_piv = gi; s_pblock = gb; }
      if (blockIdx.x == 0 && lane == 0) {
        ipiv[j0 + c] = j0 + gi;
        if (!(gv > 0.0)) atomicExch(info, 1);
        if (gv > 0.0) { loc_pmin = fmin(loc_pmin, gv); loc_pmax = fmax(loc_pmax, gv); }  // folded into piv_minmax once, after the loop
      }
    }
    __syncthreads();
    const uint64_t prow = s_piv;
    if constexpr (CLUSTER) {
      const Candidate* win = cg::this_cluster().map_shared_rank(&s_cand[buf], s_pblock);
      const double* rc = cg::this_cluster().map_shared_rank(&s_pub_rowc[buf][0], 0);  // row c always lives in CTA 0 (c < NB <= SLAB_ROWS)
      for (int cc = tid; cc < jb; cc += SLAB_ROWS) { s_row[cc] = win->row[cc]; s_rowc[cc] = rc[cc]; }
    } else {
      const Candidate* win = cand + (size_t)buf * nblk + s_pblock;
      for (int cc = tid; cc < jb; cc += SLAB_ROWS) { s_row[cc] = __ldcg(&win->row[cc]); s_rowc[cc] = __ldcg(&rowc[buf * NB + cc]); }
    }
    if (blockIdx.x == 0 && tid == 0) s_pivots[c] = prow;  // the net row permutation is composed after the column loop (off the per-column critical path)
    __syncthreads();
    if (prow != (uint64_t)c) {
      // row c (always in CTA 0) receives the pivot row; the pivot's home CTA receives the old row c
      if (blockIdx.x == 0 && tid < jb) slab[c][tid] = s_row[tid];
      const uint64_t lp = prow - (uint64_t)blockIdx.x * SLAB_ROWS;
      if (lp < (uint64_t)SLAB_ROWS && tid >= NB && tid < NB + jb) slab[lp][tid - NB] = s_rowc[tid - NB];
      __syncthreads();
    }
    const double pivot = s_row[c];
    if (valid && r > (uint64_t)c && pivot != 0.0) {
      // note: a thread that just received a swapped row (r == prow) updates the new contents
      const double l = slab[tid][c] / pivot;
      slab[tid][c] = l;
#pragma unroll 8
      for (int cc = c + 1; cc < jb; ++cc) slab[tid][cc] = fma(-l, s_row[cc], slab[tid][cc]);  // explicit DFMA (the library is built -fmad=false)
    }
    __syncthreads();
  }
  if (valid) for (int cc = 0; cc < jb; ++cc) P[r + (uint64_t)cc * lda] = slab[tid][cc];
  if (blockIdx.x == 0 && tid == 0) {
    if (loc_pmin < piv_minmax[0]) piv_minmax[0] = loc_pmin;
    if (loc_pmax > piv_minmax[1]) piv_minmax[1] = loc_pmax;
  }
  if constexpr (CLUSTER) cg::this_cluster().sync();  // nobody exits while a peer may still read its shared memory
  if (blockIdx.x == 0) {
    __syncthreads();
    // CTA 0 / warp 0: fold the jb swaps (j0+c <-> j0+pivot[c]) into the net permutation (32-lane parallel lookup of the two rows).
    // r17 ncu: doing this inside the column loop made CTA 0 the last to arrive at every cluster barrier.
    if (warp == 0) {
      for (int c = 0; c < jb; ++c) {
        const uint64_t prow = s_pivots[c];
        if (prow == (uint64_t)c) continue;
        const unsigned long long want[2] = {j0 + (uint64_t)c, j0 + prow};
        unsigned slot[2];
        for (int q = 0; q < 2; ++q) {
          unsigned found = 0xffffffffu;
          const unsigned cnt = p_n;
          for (unsigned base = 0; base < cnt; base += 32) {
            const unsigned k = base + lane;
            const unsigned hit = __ballot_sync(0xffffffffu, k < cnt && p_rows[k] == want[q]);
            if (hit) { found = base + __ffs(hit) - 1; break; }
          }
          if (found == 0xffffffffu) {
            found = cnt;
            if (lane == 0) { p_rows[cnt] = want[q]; p_cur[cnt] = want[q]; p_n = cnt + 1; }
          }
          __syncwarp();
          slot[q] = found;
        }
        if (lane == 0) { const unsigned long long t = p_cur[slot[0]]; p_cur[slot[0]] = p_cur[slot[1]]; p_cur[slot[1]] = t; }
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t k2 = 0;
      for (unsigned k = 0; k < p_n; ++k) if (p_cur[k] != p_rows[k]) { moves->dst[k2] = p_rows[k]; moves->src[k2] = p_cur[k]; ++k2; }
      moves->count = k2;
    }
  }
}

// Row interchanges of one panel, applied as a gather. The jb sequential swaps (row j0+c <-> ipiv[j0+c]) touch at most 2*jb
// distinct rows; `perm_build_kernel` (one thread) composes them into (dst_row <- src_row) moves, and `perm_apply_kernel`
// performs all moves of a column with independent loads followed by independent stores. (The first version walked the jb
// swaps sequentially per column: 64 dependent, uncoalesced round trips per thread and three launches per panel.)
// Fallback builder (used with the global-memory panel kernel, i.e. panels taller than the resident slab capacity): one warp,
// shared-memory tables.
__global__ void perm_build_kernel(const unsigned long long* __restrict__ ipiv, uint64_t j0, int jb, RowMoves* __restrict__ mv) {
  __shared__ unsigned long long rows[2 * NB], cur[2 * NB];
  if (threadIdx.x != 0) return;
  uint32_t n = 0;
  auto slot = [&](unsigned long long r) -> uint32_t {
    for (uint32_t k = 0; k < n; ++k) if (rows[k] == r) return k;
    rows[n] = r; cur[n] = r;
    return n++;
  };
  for (int c = 0; c < jb; ++c) {
    const unsigned long long r0 = j0 + c, r1 = ipiv[j0 + c];
    if (r0 == r1) continue;
    const uint32_t a = slot(r0), b = slot(r1);
    const unsigned long long t = cur[a]; cur[a] = cur[b]; cur[b] = t;
  }
  uint32_t m = 0;
  for (uint32_t k = 0; k < n; ++k) if (cur[k] != rows[k]) { mv->dst[m] = rows[k]; mv->src[m] = cur[k]; ++m; }
  mv->count = m;
}
// Columns [0, ncols) of the logical range; columns >= skip_from are shifted by skip_len (to jump over the panel itself).
__global__ void __launch_bounds__(2 * NB) perm_apply_kernel(double* __restrict__ M, uint64_t ld, uint64_t ncols, uint64_t skip_from, uint64_t skip_len,
                                                            const RowMoves* __restrict__ mv) {
  const uint32_t cnt = mv->count;
  if (cnt == 0) return;
  const bool active = threadIdx.x < cnt;
  const unsigned long long d = active ? mv->dst[threadIdx.x] : 0, sr = active ? mv->src[threadIdx.x] : 0;
  for (uint64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    const uint64_t col = c < skip_from ? c : c + skip_len;
    double* colp = M + col * ld;
    const double v = active ? colp[sr] : 0.0;
    __syncthreads();  // all loads of this column before any store (moves may chain)
    if (active) colp[d] = v;
    __syncthreads();
  }
}

// Solve T * X = Bm in place for a jb x jb triangular block T (ldt) and jb x ncols right-hand sides Bm (ldb).
// MODE 1 (TRSM_LOWER_UNIT): forward substitution with implicit unit diagonal (the L of the LU); MODE 0 (TRSM_UPPER): backward
// substitution with the stored diagonal; MODE 2 (TRSM_LOWER): forward substitution with the stored diagonal (linsolve's
// general lower-triangular systems). One CTA handles 32 columns; T and the column tile live in shared memory.
constexpr int TRSM_UPPER = 0, TRSM_LOWER_UNIT = 1, TRSM_LOWER = 2;
// r06 launch list: the first version (columns in shared memory, 8 row-parts per column, two __syncthreads per substitution step)
// took 22 us per 64x64 block whatever the number of right-hand sides, ~240 launches per 4096 solve. Here a warp owns 4 columns
// and keeps them in REGISTERS (lane l holds rows l and l+32): a substitution step is one shuffle broadcast of x_i per column, two
// conflict-free shared loads of T's column i and two DFMAs per column -- no block-wide barrier inside the 64-step chain.
template <int MODE>
__global__ void __launch_bounds__(256) trsm_block_kernel(const double* __restrict__ T, uint64_t ldt, int jb, double* __restrict__ Bm, uint64_t ldb, uint64_t ncols) {
  extern __shared__ double trsm_smem[];
  double (*sT)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(trsm_smem);  // [row][col], padded: lanes walking rows hit distinct banks
  for (int i = threadIdx.x; i < NB * NB; i += 256) {
    const int r = i % NB, c = i / NB;
    sT[r][c] = (r < jb && c < jb) ? T[r + (uint64_t)c * ldt] : (r == c ? 1.0 : 0.0);  // identity padding: rows >= jb never change anything
  }
  __syncthreads();
  // reciprocals of the diagonal, computed once and in parallel: a division inside the 64-step dependent chain cost more than the
  // rest of a step (r17: the upper solves ran 25 us, the unit-diagonal ones 10 us)
  __shared__ double s_rdiag[NB];
  if (MODE != TRSM_LOWER_UNIT && threadIdx.x < NB) s_rdiag[threadIdx.x] = 1.0 / sT[threadIdx.x][threadIdx.x];
  __syncthreads();
  constexpr int CW = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t c0 = (uint64_t)blockIdx.x * 32 + (uint64_t)warp * CW;
  double b0[CW], b1[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    const bool ok = c0 + c < ncols;
    b0[c] = (ok && lane < jb) ? Bm[lane + (c0 + c) * ldb] : 0.0;
    b1[c] = (ok && lane + 32 < jb) ? Bm[lane + 32 + (c0 + c) * ldb] : 0.0;
  }
  if (MODE == TRSM_LOWER_UNIT || MODE == TRSM_LOWER) {
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const double l0 = sT[lane][i], l1 = sT[lane + 32][i];
      const double d = MODE == TRSM_LOWER ? s_rdiag[i] : 1.0;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double xi = __shfl_sync(0xffffffffu, b0[c], i);
        if (MODE == TRSM_LOWER) { xi = xi * d; if (lane == i) b0[c] = xi; }
        if (lane > i) b0[c] -= l0 * xi;
        b1[c] -= l1 * xi;
      }
    }
#pragma unroll 4
    for (int i = 32; i < NB; ++i) {
      const double l1 = sT[lane + 32][i];
      const double d = MODE == TRSM_LOWER ? s_rdiag[i] : 1.0;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double xi = __shfl_sync(0xffffffffu, b1[c], i - 32);
        if (MODE == TRSM_LOWER) { xi = xi * d; if (lane == i - 32) b1[c] = xi; }
        if (lane + 32 > i) b1[c] -= l1 * xi;
      }
    }
  } else {
#pragma unroll 4
    for (int i = NB - 1; i >= 32; --i) {
      const double u0 = sT[lane][i], u1 = sT[lane + 32][i];
      const double d = s_rdiag[i];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double xi = __shfl_sync(0xffffffffu, b1[c], i - 32);
        xi = xi * d;
        if (lane == i - 32) b1[c] = xi;
        if (lane + 32 < i) b1[c] -= u1 * xi;
        b0[c] -= u0 * xi;
      }
    }
#pragma unroll 4
    for (int i = 31; i >= 0; --i) {
      const double u0 = sT[lane][i];
      const double d = s_rdiag[i];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        double xi = __shfl_sync(0xffffffffu, b0[c], i);
        xi = xi * d;
        if (lane == i) b0[c] = xi;
        if (lane < i) b0[c] -= u0 * xi;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    if (c0 + c < ncols) {
      if (lane < jb) Bm[lane + (c0 + c) * ldb] = b0[c];
      if (lane + 32 < jb) Bm[lane + 32 + (c0 + c) * ldb] = b1[c];
    }
  }
}

// dst = src while tracking max |x| (as an order-preserving bit pattern) and non-finite entries: 4 independent loads in flight per
// thread, warp shuffle fold, one atomic per warp. (Round 2's stand-alone scan issued one atomicMax per THREAD on a single address
// and took 228 us for a 4096 x 4096 matrix: r44 launch list.)
__global__ void __launch_bounds__(256) copy_absmax_kernel(const double* __restrict__ src, double* __restrict__ dst, uint64_t n,
                                                          unsigned long long* __restrict__ out, int* __restrict__ nonfinite) {
  double mx = 0.0;
  bool bad = false;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dst[i + u * stride] = v[u];
      const double av = fabs(v[u]);
      bad |= !(av <= 1.7976931348623157e308);
      mx = fmax(mx, av);
    }
  }
  for (; i < n; i += stride) {
    const double v = src[i];
    dst[i] = v;
    const double av = fabs(v);
    bad |= !(av <= 1.7976931348623157e308);
    mx = fmax(mx, av);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(nonfinite, 1);
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// min / max of |diag(T)| (forward/backward_substitution_real track exactly these, linsolve.rs:769-833) as IEEE bit patterns
__global__ void diag_minmax_kernel(const double* __restrict__ T, uint64_t n, unsigned long long* __restrict__ mm) {
  double mn = 1.7976931348623157e308 * 2.0, mx = 0.0;  // +inf, 0
  bool nan = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = fabs(T[i + i * n]);
    nan |= d != d;
    mn = fmin(mn, d);
    mx = fmax(mx, d);
  }
  atomicMin(&mm[0], (unsigned long long)__double_as_longlong(mn));
  atomicMax(&mm[1], (unsigned long long)__double_as_longlong(mx));
  if (nan) atomicExch(&mm[2], 1ull);
}

}  // namespace

}  // namespace rm

using namespace rm;

// linsolve (lib.rs:2422-2476; host semantics: builtins/math/linalg/solve/linsolve.rs:691-726). Device paths, mirroring what
// the wgpu provider keeps on the device (backend/wgpu/provider/ops/solve.rs:869-950):
//  * LT / UT (exactly one): blocked forward / backward substitution (64x64 diagonal blocks in shared memory + DMMA updates).
//    TRANSA solves with A' (one device transpose; the triangle hint flips, linsolve.rs:698-705). A zero diagonal entry is
//    the host's "singular to working precision" error; rcond = min|diag| / max|diag| (diagonal_rcond, common/linalg.rs:232)
//    is always returned, and enforced against opts.rcond like enforce_rcond (:1000-1009).
//  * no structure hint, SYM or POSDEF: square systems go through the LU of rm_mldivide when the caller does not need the
//    reciprocal condition number (the host derives it from singular values, which an LU cannot reproduce): rcond = NaN, as the
//    wgpu provider reports when it was not asked for one. Otherwise RM_UNSUPPORTED -> host fallback.
//  * RECT: RM_UNSUPPORTED.
RM_EXPORT rm_status rm_linsolve(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, const rm_linsolve_options* opt, rm_handle* solution,
                                double* reciprocal_condition) {
  RM_REQUIRE(p && lhs && rhs && opt && solution && reciprocal_condition, RM_INVALID_ARG, "linsolve: bad arguments");
  RM_REQUIRE(lhs->rank <= 2 && rhs->rank <= 2, RM_ERROR, "linsolve: inputs must be 2-D matrices or vectors");
  RM_REQUIRE(p->precision == RM_F64, RM_UNSUPPORTED, "linsolve: f32 storage not supported by provider");
  RM_REQUIRE(!opt->rectangular, RM_UNSUPPORTED, "linsolve: RECT (least-squares) solve not supported by provider");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_linsolve);
  *reciprocal_condition = NAN;
  const bool needs_rcond = opt->need_rcond || opt->has_rcond;
  auto dim = [](const rm_handle* h, int d) -> uint64_t { return d < (int)h->rank ? h->shape[d] : 1; };
  const uint64_t arows = opt->transposed ? dim(lhs, 1) : dim(lhs, 0), acols = opt->transposed ? dim(lhs, 0) : dim(lhs, 1);
  const uint64_t brows = dim(rhs, 0), nrhs = dim(rhs, 1);
  const bool tri = (opt->lower != 0) != (opt->upper != 0) && !opt->symmetric && !opt->posdef;
  rm_handle a_eff = *lhs;
  bool own_a = false;
  auto drop_a = [&]() { if (own_a) rm_free(p, &a_eff); };
  if (!tri) {
    RM_REQUIRE(!needs_rcond, RM_UNSUPPORTED, "linsolve: singular-value reciprocal condition estimate not supported by provider");
    RM_REQUIRE(arows == acols, RM_UNSUPPORTED, "linsolve: non-square general solve not supported by provider");
    RM_REQUIRE(brows == arows, RM_ERROR, "Matrix dimensions must agree.");  // normalize_rhs_tensor, linsolve.rs:972-984
    if (opt->transposed) { RM_TRY(rm_transpose(p, lhs, &a_eff)); own_a = true; }
    rm_handle a2 = a_eff, b2 = *rhs;
    a2.rank = 2; a2.shape[0] = arows; a2.shape[1] = acols;
    b2.rank = 2; b2.shape[0] = brows; b2.shape[1] = nrhs;
    rm_status st = rm_mldivide(p, &a2, &b2, solution);
    std::string msg = st == RM_OK ? "" : last_error();
    drop_a();
    if (st != RM_OK) set_error("%s", msg.c_str());
    return st;
  }
  RM_REQUIRE(arows == acols, RM_ERROR, "linsolve: triangular solves require a square coefficient matrix.");  // ensure_square, :1057-1065
  const uint64_t n = arows;
  RM_REQUIRE(brows == n, RM_ERROR, "Matrix dimensions must agree.");
  uint64_t oshape[2] = {n, nrhs};
  if (n == 0 || nrhs == 0) return rm_zeros(p, oshape, 2, solution);
  const bool eff_lower = opt->transposed ? opt->upper != 0 : opt->lower != 0;  // the hint describes A; A' flips it
  if (opt->transposed) { RM_TRY(rm_transpose(p, lhs, &a_eff)); own_a = true; }
  void *pa, *pb, *px;
  rm_status st = resolve(p, &a_eff, &pa, nullptr);
  if (st == RM_OK) st = resolve(p, rhs, &pb, nullptr);
  if (st != RM_OK) { std::string m = last_error(); drop_a(); set_error("%s", m.c_str()); return st; }
  cudaStream_t stream = p->stream;
  unsigned long long* mm = nullptr;
  auto bail = [&](rm_status code, const char* msg) { if (mm) cudaFreeAsync(mm, stream); drop_a(); return fail(code, "%s", msg); };
  if (cudaMallocAsync((void**)&mm, 24, stream) != cudaSuccess) { cudaGetLastError(); return bail(RM_OOM, "linsolve: allocation failed"); }
  const unsigned long long init[3] = {0x7ff0000000000000ull, 0ull, 0ull};
  cudaMemcpyAsync(mm, init, 24, cudaMemcpyHostToDevice, stream);
  diag_minmax_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, stream>>>((const double*)pa, n, mm);
  count_launch(p);
  unsigned long long h_mm[3];
  cudaMemcpyAsync(h_mm, mm, 24, cudaMemcpyDeviceToHost, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess) { cudaGetLastError(); return bail(RM_ERROR, "linsolve: device error while scanning the diagonal"); }
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  double dmin, dmax;
  memcpy(&dmin, &h_mm[0], 8);
  memcpy(&dmax, &h_mm[1], 8);
  if (dmin == 0.0) return bail(RM_ERROR, "linsolve: matrix is singular to working precision.");
  const double rcond = h_mm[2] ? NAN : (dmax == 0.0 ? 0.0 : dmin / dmax);  // diagonal_rcond
  if (opt->has_rcond && rcond < opt->rcond) return bail(RM_ERROR, "linsolve: matrix is singular to working precision.");  // enforce_rcond
  cudaFreeAsync(mm, stream);
  mm = nullptr;
  st = alloc_tensor(p, oshape, 2, solution, &px);
  if (st != RM_OK) { std::string m = last_error(); drop_a(); set_error("%s", m.c_str()); return st; }
  const double* T = (const double*)pa;
  double* X = (double*)px;
  cudaMemcpyAsync(X, pb, n * nrhs * 8, cudaMemcpyDeviceToDevice, stream);
  constexpr size_t TRSM_SMEM = (size_t)(NB + 32) * (NB + 1) * sizeof(double);
  cudaFuncSetAttribute(trsm_block_kernel<TRSM_LOWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  cudaFuncSetAttribute(trsm_block_kernel<TRSM_UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  const unsigned tgrid = (unsigned)((nrhs + 31) / 32);
  if (eff_lower) {
    for (uint64_t j0 = 0; j0 < n && st == RM_OK; j0 += NB) {
      const int jb = (int)std::min<uint64_t>(NB, n - j0);
      trsm_block_kernel<TRSM_LOWER><<<tgrid, 256, TRSM_SMEM, stream>>>(T + j0 + j0 * n, n, jb, X + j0, n, nrhs);
      count_launch(p);
      const uint64_t rest = n - j0 - jb;
      if (rest > 0) st = dgemm_sub_strided(p, T + (j0 + jb) + j0 * n, n, X + j0, n, X + j0 + jb, n, rest, nrhs, (uint64_t)jb);
    }
  } else {
    for (uint64_t jend = n; jend > 0 && st == RM_OK;) {
      const uint64_t j0 = ((jend - 1) / NB) * NB;
      const int jb = (int)(jend - j0);
      trsm_block_kernel<TRSM_UPPER><<<tgrid, 256, TRSM_SMEM, stream>>>(T + j0 + j0 * n, n, jb, X + j0, n, nrhs);
      count_launch(p);
      if (j0 > 0) st = dgemm_sub_strided(p, T + j0 * n, n, X + j0, n, X, n, j0, nrhs, (uint64_t)jb);
      jend = j0;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (st == RM_OK && e != cudaSuccess) st = fail(RM_ERROR, "linsolve launch failed: %s", cudaGetErrorString(e));
  std::string msg = st == RM_OK ? "" : last_error();
  drop_a();
  if (st != RM_OK) { rm_free(p, solution); set_error("%s", msg.c_str()); return st; }
  *reciprocal_condition = rcond;
  return RM_OK;
}

RM_EXPORT rm_status rm_mldivide(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out) {
  RM_REQUIRE(p && a && b && out, RM_INVALID_ARG, "mldivide: bad arguments");
  RM_REQUIRE(a->rank == 2 && b->rank == 2, RM_ERROR, "mldivide: inputs must be 2-D matrices");
  RM_REQUIRE(p->precision == RM_F64, RM_UNSUPPORTED, "mldivide: f32 storage not supported by provider");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_mldivide);
  void *pa, *pb;
  RM_TRY(resolve(p, a, &pa, nullptr));
  RM_TRY(resolve(p, b, &pb, nullptr));
  const uint64_t n = a->shape[0], nrhs = b->shape[1];
  if (a->shape[0] == 1 && a->shape[1] == 1) {  // scalar divisor: rhs * (1/a)  (mldivide.rs:321-325)
    double av;
    RM_TRY(rm_read_scalar(p, a, 0, &av));
    return rm_scalar_op_apply(p, RM_SC_MUL, b, 1.0 / av, out);
  }
  RM_REQUIRE(a->shape[0] == b->shape[0], RM_ERROR, "mldivide: left and right operands must have the same number of rows (%llu vs %llu)",
             (unsigned long long)a->shape[0], (unsigned long long)b->shape[0]);
  RM_REQUIRE(a->shape[0] == a->shape[1], RM_UNSUPPORTED, "mldivide: least-squares (non-square) solve not supported by provider");
  uint64_t oshape[2] = {n, nrhs};
  if (n == 0 || nrhs == 0) return rm_zeros(p, oshape, 2, out);

  cudaStream_t st = p->stream;
  double* LU = nullptr;
  unsigned long long* ipiv = nullptr;
  PivotEntry* scratch = nullptr;
  int* info = nullptr;          // [0] singular, [1] non-finite input
  double* pivmm = nullptr;      // [0] min |pivot|, [1] max |pivot|
  unsigned long long* amax = nullptr;
  void* px = nullptr;
  bool have_out = false;
  RowMoves* moves = nullptr;
  Candidate* cand = nullptr;
  double* rowbuf = nullptr;  // [2][NB]: pivot row + displaced row exchanged between CTAs by the slab panel kernel
  auto cleanup = [&](bool drop_out) {
    if (LU) cudaFreeAsync(LU, st);
    if (ipiv) cudaFreeAsync(ipiv, st);
    if (scratch) cudaFreeAsync(scratch, st);
    if (info) cudaFreeAsync(info, st);
    if (pivmm) cudaFreeAsync(pivmm, st);
    if (amax) cudaFreeAsync(amax, st);
    if (rowbuf) cudaFreeAsync(rowbuf, st);
    if (moves) cudaFreeAsync(moves, st);
    if (cand) cudaFreeAsync(cand, st);
    if (drop_out && have_out) rm_free(p, out);
  };
#define SV_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); cleanup(true); return fail(_e == cudaErrorMemoryAllocation ? RM_OOM : RM_ERROR, "%s failed: %s", #expr, cudaGetErrorString(_e)); } } while (0)
#define SV_TRY(expr) do { rm_status _s = (expr); if (_s != RM_OK) { std::string _m = last_error(); cleanup(true); set_error("%s", _m.c_str()); return _s; } } while (0)

  constexpr size_t TRSM_SMEM = (size_t)(NB + 32) * (NB + 1) * sizeof(double);
  cudaFuncSetAttribute(trsm_block_kernel<TRSM_LOWER_UNIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  cudaFuncSetAttribute(trsm_block_kernel<TRSM_UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM);
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->ordinal);
  RM_REQUIRE(coop, RM_UNSUPPORTED, "mldivide: cooperative launch not supported on this device");
  int max_blocks_per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, lu_panel_kernel, 256, 0);
  const unsigned max_grid = (unsigned)std::max(1, max_blocks_per_sm) * (unsigned)p->prop.multiProcessorCount;

  constexpr size_t SLAB_SMEM = (size_t)SLAB_ROWS * (NB + 1) * sizeof(double);
  cudaFuncSetAttribute(lu_panel_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SLAB_SMEM);
  cudaFuncSetAttribute(lu_panel_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SLAB_SMEM);
  cudaFuncSetAttribute(lu_panel_smem_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  // largest power-of-two cluster (<= 16 CTAs = 4096 panel rows) the device can co-schedule with this footprint
  unsigned cluster_max = 0;
  if (!getenv("RUNMAT_B200_LU_NO_CLUSTER")) {
    for (unsigned cs = 16; cs >= 1 && !cluster_max; cs >>= 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(SLAB_ROWS); cfg.dynamicSmemBytes = SLAB_SMEM; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, lu_panel_smem_kernel<true>, &cfg) == cudaSuccess && nc >= 1) cluster_max = cs;
    }
    cudaGetLastError();
  }
  // register-resident push kernel (lu_panel_push.h): the default whenever a full 64-column panel fits one cluster
  unsigned push_cluster_max = 0;
  if (cluster_max && !getenv("RUNMAT_B200_LU_PANEL_V1")) {
    cudaFuncSetAttribute(lupush::lu_panel_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lupush::Smem));
    cudaFuncSetAttribute(lupush::lu_panel_push_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (unsigned cs = 16; cs >= 1 && !push_cluster_max; cs >>= 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs); cfg.blockDim = dim3(lupush::ROWS); cfg.dynamicSmemBytes = sizeof(lupush::Smem); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, lupush::lu_panel_push_kernel, &cfg) == cudaSuccess && nc >= 1) push_cluster_max = cs;
    }
    cudaGetLastError();
  }
  int slab_blocks_per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&slab_blocks_per_sm, lu_panel_smem_kernel<false>, SLAB_ROWS, SLAB_SMEM);
  const unsigned slab_max_grid = (unsigned)std::max(0, slab_blocks_per_sm) * (unsigned)p->prop.multiProcessorCount;
  SV_CUDA(cudaMallocAsync((void**)&rowbuf, 2 * NB * 8, st));
  SV_CUDA(cudaMallocAsync((void**)&cand, (size_t)2 * std::max(slab_max_grid, 1u) * sizeof(Candidate), st));
  SV_CUDA(cudaMallocAsync((void**)&moves, sizeof(RowMoves) * ((n + NB - 1) / NB), st));  // one net permutation per panel (applied by both streams)
  // Augmented system [A | B] in one column-major buffer (ld = n): the row interchanges, the U12 solves and the trailing updates
  // of the factorisation then carry the right-hand sides along, i.e. the forward substitution L y = P b costs no launches of its
  // own (r06: 64 triangular solves + 63 skinny GEMMs = 3.3 ms of the 32 ms at n = 4096).
  const uint64_t ntot = n + nrhs;
  SV_CUDA(cudaMallocAsync((void**)&LU, n * ntot * 8, st));
  SV_CUDA(cudaMallocAsync((void**)&ipiv, n * 8, st));
  SV_CUDA(cudaMallocAsync((void**)&scratch, (size_t)max_grid * sizeof(PivotEntry), st));
  SV_CUDA(cudaMallocAsync((void**)&info, 8, st));
  SV_CUDA(cudaMallocAsync((void**)&pivmm, 16, st));
  SV_CUDA(cudaMallocAsync((void**)&amax, 8, st));
  SV_CUDA(cudaMemsetAsync(info, 0, 8, st));
  SV_CUDA(cudaMemsetAsync(amax, 0, 8, st));
  const double mm0[2] = {1.7976931348623157e308, 0.0};
  SV_CUDA(cudaMemcpyAsync(pivmm, mm0, 16, cudaMemcpyHostToDevice, st));
  // A -> LU copy fused with the max |a| / non-finite scan the conditioning gate needs (one pass over A instead of a copy plus a scan)
  copy_absmax_kernel<<<(unsigned)p->prop.multiProcessorCount * 4, 256, 0, st>>>((const double*)pa, LU, n * n, amax, info + 1);
  count_launch(p);
  SV_TRY(alloc_tensor(p, oshape, 2, out, &px));
  have_out = true;
  double* X = LU + n * n;  // the right-hand sides ride in the augmented columns; copied to the result tensor at the end
  SV_CUDA(cudaMemcpyAsync(X, pb, n * nrhs * 8, cudaMemcpyDeviceToDevice, st));

  // Two-level blocking with look-ahead on two streams.
  //  * 64-wide panels are factored and applied only inside the current NBO-wide outer block (stream A = the provider stream: the
  //    latency-bound critical path: cluster panel kernel, in-block row interchanges, in-block triangular solve + k=64 update);
  //  * everything to the right of the block -- row interchanges, U12 <- L11^-1 A12 (the right-hand sides ride along as extra
  //    columns), the rank-NBO trailing update -- runs on stream B. B updates the NEXT block's columns first and signals A, so the
  //    panels of block K+1 overlap the bulk of block K's trailing update (r06: panels 16 ms, everything else 16 ms, serialised).
  //  * Row interchanges are never applied to the columns LEFT of the current block: those hold finished L factors that nothing
  //    reads again (the forward substitution is folded into the factorisation, the back substitution only needs U), and leaving
  //    them alone is what makes A(K+1) and B(K) touch disjoint memory.
  uint64_t NBO = 256;
  if (const char* e = getenv("RUNMAT_B200_LU_OUTER")) { const long v = atol(e); if (v >= NB && v % NB == 0) NBO = (uint64_t)v; }
  const bool lookahead = !getenv("RUNMAT_B200_LU_NO_LOOKAHEAD");
  if (lookahead && !p->aux_stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    SV_CUDA(cudaStreamCreateWithPriority(&p->aux_stream, cudaStreamNonBlocking, lo));  // lowest priority: panel CTAs go first
  }
  cudaStream_t sb = lookahead ? p->aux_stream : st;
  const uint64_t nblocks = (n + NBO - 1) / NBO;
  std::vector<cudaEvent_t> evE(nblocks, nullptr), evF(nblocks + 1, nullptr);
  auto destroy_events = [&] { for (auto e : evE) if (e) cudaEventDestroy(e); for (auto e : evF) if (e) cudaEventDestroy(e); };
  if (lookahead) {
    for (auto& e : evE) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& e : evF) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventRecord(evF[0], st);  // B starts after the set-up copies on A
    cudaStreamWaitEvent(sb, evF[0], 0);
  }
  rm_status loop_status = RM_OK;
  cudaError_t loop_err = cudaSuccess;
  for (uint64_t K = 0; K < nblocks && loop_status == RM_OK && loop_err == cudaSuccess; ++K) {
    const uint64_t J0 = K * NBO, Jend = std::min<uint64_t>(J0 + NBO, n);
    if (lookahead && K > 0) cudaStreamWaitEvent(st, evF[K], 0);  // block K's columns carry every earlier update
    for (uint64_t j0 = J0; j0 < Jend && loop_status == RM_OK && loop_err == cudaSuccess; j0 += NB) {
      int jb = (int)std::min<uint64_t>(NB, n - j0);
      const uint64_t m = n - j0;
      RowMoves* mv = moves + j0 / NB;
      unsigned grid = (unsigned)std::min<uint64_t>((m + 255) / 256, max_grid);
      grid = std::max(grid, 1u);
      uint64_t lda = n, nn = n, jj = j0;
      const unsigned slab_grid = (unsigned)((m + SLAB_ROWS - 1) / SLAB_ROWS);
      unsigned cl = 1;
      while (cl < slab_grid) cl <<= 1;
      if (jb == NB && cl <= push_cluster_max && !getenv("RUNMAT_B200_LU_GLOBAL_PANEL")) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cl); cfg.blockDim = dim3(lupush::ROWS); cfg.dynamicSmemBytes = sizeof(lupush::Smem); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        loop_err = cudaLaunchKernelEx(&cfg, lupush::lu_panel_push_kernel, LU, lda, nn, jj, ipiv, info, pivmm, mv);
      } else if (cl <= cluster_max && !getenv("RUNMAT_B200_LU_GLOBAL_PANEL")) {
        // whole panel inside one thread-block cluster (grid rounded up to a power of two; surplus CTAs own no rows)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cl); cfg.blockDim = dim3(SLAB_ROWS); cfg.dynamicSmemBytes = SLAB_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        loop_err = cudaLaunchKernelEx(&cfg, lu_panel_smem_kernel<true>, LU, lda, nn, jj, jb, ipiv, cand, rowbuf, info, pivmm, mv);
      } else if (slab_grid <= slab_max_grid && slab_grid <= max_grid && !getenv("RUNMAT_B200_LU_GLOBAL_PANEL")) {
        void* args[] = {&LU, &lda, &nn, &jj, &jb, &ipiv, &cand, &rowbuf, &info, &pivmm, &mv};
        loop_err = cudaLaunchCooperativeKernel((void*)lu_panel_smem_kernel<false>, dim3(slab_grid), dim3(SLAB_ROWS), args, SLAB_SMEM, st);
      } else {
        void* args[] = {&LU, &lda, &nn, &jj, &jb, &ipiv, &scratch, &info, &pivmm};
        loop_err = cudaLaunchCooperativeKernel((void*)lu_panel_kernel, dim3(grid), dim3(256), args, 0, st);
        perm_build_kernel<<<1, 32, 0, st>>>(ipiv, j0, jb, mv);
      }
      if (loop_err != cudaSuccess) break;
      // row interchanges on the other columns of this outer block
      const uint64_t inblk = Jend - J0 - jb;
      if (inblk > 0) perm_apply_kernel<<<(unsigned)std::min<uint64_t>(inblk, 4096), 2 * NB, 0, st>>>(LU + J0 * n, n, inblk, j0 - J0, (uint64_t)jb, mv);
      count_launch(p, 2);
      const uint64_t rest = n - j0 - jb;
      const uint64_t rest_in = Jend - (j0 + jb);  // columns of the outer block still to factor
      if (rest_in > 0) {
        // inside the outer block: A12 <- L11^-1 A12 ; A22 -= A21 * A12
        trsm_block_kernel<TRSM_LOWER_UNIT><<<(unsigned)((rest_in + 31) / 32), 256, TRSM_SMEM, st>>>(LU + j0 + j0 * n, n, jb, LU + j0 + (j0 + jb) * n, n, rest_in);
        count_launch(p);
        loop_status = dgemm_sub_strided(p, LU + (j0 + jb) + j0 * n, n, LU + j0 + (j0 + jb) * n, n, LU + (j0 + jb) + (j0 + jb) * n, n, rest, rest_in, (uint64_t)jb, st);
      }
    }
    if (loop_status != RM_OK || loop_err != cudaSuccess) break;
    // ---- right of the block (stream B): interchanges, U12, trailing update ----
    if (lookahead) { cudaEventRecord(evE[K], st); cudaStreamWaitEvent(sb, evE[K], 0); }
    const uint64_t right = ntot - Jend;
    if (right > 0) {
      // The next outer block's columns go FIRST through the whole chain (interchanges, U12 solves, their trailing update) and
      // release stream A; the remaining columns follow. (r44 launch list: with one pass over all right-hand columns A waited
      // ~170 us per outer block for 4 interchange + 4 solve + 3 update launches over up to 3900 columns before its panels could start.)
      const uint64_t next_cols = n > Jend ? std::min<uint64_t>(NBO, n - Jend) : 0;
      auto right_pass = [&](uint64_t c0, uint64_t nc) {  // columns [Jend + c0, Jend + c0 + nc) of the augmented matrix
        if (nc == 0 || loop_status != RM_OK) return;
        double* Mc = LU + (Jend + c0) * n;
        for (uint64_t j0 = J0; j0 < Jend; j0 += NB)
          perm_apply_kernel<<<(unsigned)std::min<uint64_t>(nc, 4096), 2 * NB, 0, sb>>>(Mc, n, nc, nc, 0, moves + j0 / NB);
        count_launch(p, (Jend - J0 + NB - 1) / NB);
        for (uint64_t i0 = J0; i0 < Jend && loop_status == RM_OK; i0 += NB) {
          const int ib = (int)std::min<uint64_t>(NB, Jend - i0);
          trsm_block_kernel<TRSM_LOWER_UNIT><<<(unsigned)((nc + 31) / 32), 256, TRSM_SMEM, sb>>>(LU + i0 + i0 * n, n, ib, Mc + i0, n, nc);
          count_launch(p);
          const uint64_t below = Jend - (i0 + ib);
          if (below > 0) loop_status = dgemm_sub_strided(p, LU + (i0 + ib) + i0 * n, n, Mc + i0, n, Mc + (i0 + ib), n, below, nc, (uint64_t)ib, sb);
        }
        if (n > Jend && loop_status == RM_OK)
          loop_status = dgemm_sub_strided(p, LU + Jend + J0 * n, n, Mc + J0, n, Mc + Jend, n, n - Jend, nc, Jend - J0, sb);
      };
      right_pass(0, next_cols);
      if (lookahead && next_cols > 0) cudaEventRecord(evF[K + 1], sb);
      right_pass(next_cols, right - next_cols);
    }
  }
  if (lookahead) {
    // A continues (conditioning gate, back substitution) once B has drained
    cudaEventRecord(evF[nblocks], sb);
    cudaStreamWaitEvent(st, evF[nblocks], 0);
    destroy_events();  // destruction is deferred by the runtime until the recorded work has completed
  }
  if (loop_err != cudaSuccess) { cudaGetLastError(); cleanup(true); return fail(RM_ERROR, "mldivide: panel launch failed: %s", cudaGetErrorString(loop_err)); }
  SV_TRY(loop_status);
  SV_CUDA(cudaGetLastError());

  // conditioning / singularity gate (one small D2H): fall back to the host SVD path when LU is not trustworthy
  int h_info[2];
  double h_mm[2];
  unsigned long long h_amax;
  SV_CUDA(cudaMemcpyAsync(h_info, info, 8, cudaMemcpyDeviceToHost, st));
  SV_CUDA(cudaMemcpyAsync(h_mm, pivmm, 16, cudaMemcpyDeviceToHost, st));
  SV_CUDA(cudaMemcpyAsync(&h_amax, amax, 8, cudaMemcpyDeviceToHost, st));
  SV_CUDA(cudaStreamSynchronize(st));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  double amaxv;
  memcpy(&amaxv, &h_amax, 8);
  if (h_info[0] == lupush::INFO_TIMEOUT) { cleanup(true); return fail(RM_ERROR, "mldivide: the panel kernel's cluster exchange timed out (device protocol failure)"); }
  if (h_info[1]) { cleanup(true); return fail(RM_UNSUPPORTED, "mldivide: non-finite input not supported by provider"); }
  if (h_info[0] || !(h_mm[0] > (double)n * 2.220446049250313e-16 * amaxv)) {
    cleanup(true);
    return fail(RM_UNSUPPORTED, "mldivide: matrix is singular or badly conditioned for LU (min pivot %.3e, max |A| %.3e); not supported by provider", h_info[0] ? 0.0 : h_mm[0], amaxv);
  }

  // backward substitution: U x = y
  for (uint64_t jend = n; jend > 0;) {
    const uint64_t j0 = ((jend - 1) / NB) * NB;
    const int jb = (int)(jend - j0);
    trsm_block_kernel<TRSM_UPPER><<<(unsigned)((nrhs + 31) / 32), 256, TRSM_SMEM, st>>>(LU + j0 + j0 * n, n, jb, X + j0, n, nrhs);
    count_launch(p);
    if (j0 > 0) SV_TRY(dgemm_sub_strided(p, LU + j0 * n, n, X + j0, n, X, n, j0, nrhs, (uint64_t)jb));
    jend = j0;
  }
  SV_CUDA(cudaMemcpyAsync(px, X, n * nrhs * 8, cudaMemcpyDeviceToDevice, st));
  SV_CUDA(cudaGetLastError());
  cleanup(false);
#undef SV_CUDA
#undef SV_TRY
  record_launch(p, "mldivide_lu", {{"n", n}, {"nrhs", nrhs}}, {{"panel", NB}, {"outer", NBO}, {"lookahead", lookahead ? 1ull : 0ull}});
  return RM_OK;
}

// mrdivide (lib.rs:2484): lhs / rhs = (rhs' \ lhs')'
RM_EXPORT rm_status rm_mrdivide(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, rm_handle* out) {
  RM_REQUIRE(p && lhs && rhs && out, RM_INVALID_ARG, "mrdivide: bad arguments");
  RM_REQUIRE(lhs->rank == 2 && rhs->rank == 2, RM_ERROR, "mrdivide: inputs must be 2-D matrices");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_mrdivide);
  if (rhs->shape[0] == 1 && rhs->shape[1] == 1) {
    double rv;
    RM_TRY(rm_read_scalar(p, rhs, 0, &rv));
    return rm_scalar_op_apply(p, RM_SC_MUL, lhs, 1.0 / rv, out);
  }
  RM_REQUIRE(lhs->shape[1] == rhs->shape[1], RM_ERROR, "mrdivide: left and right operands must have the same number of columns (%llu vs %llu)",
             (unsigned long long)lhs->shape[1], (unsigned long long)rhs->shape[1]);
  rm_handle lt, rt, xt;
  RM_TRY(rm_transpose(p, lhs, &lt));
  rm_status st = rm_transpose(p, rhs, &rt);
  if (st != RM_OK) { std::string m = last_error(); rm_free(p, &lt); set_error("%s", m.c_str()); return st; }
  st = rm_mldivide(p, &rt, &lt, &xt);
  std::string msg = st == RM_OK ? "" : last_error();
  rm_free(p, &lt);
  rm_free(p, &rt);
  if (st != RM_OK) { set_error("%s", msg.c_str()); return st; }
  st = rm_transpose(p, &xt, out);
  msg = st == RM_OK ? "" : last_error();
  rm_free(p, &xt);
  if (st != RM_OK) set_error("%s", msg.c_str());
  return st;
}
