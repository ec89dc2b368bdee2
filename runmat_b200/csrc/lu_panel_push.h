// lu_panel_push.h — LU panel factorisation (64 columns, partial pivoting) for one thread-block cluster; used by solve.cu (mldivide,
// row a9 / SURVEY §8f #1) and by the bring-up harness scripts/lu_dev/panel_test.cu.
//
// Why a second cluster kernel. lu_panel_smem_kernel<true> (solve.cu) spends 3.9 us per column = 16 of mldivide's 23 ms at n = 4096:
//   * its per-column exchange is `st.release.cluster` / `ld.acquire.cluster` on generic addresses, which ptxas lowers to
//     MEMBAR.ALL.GPU + ST.E.STRONG.GPU and LD.E.STRONG.GPU + CCTL.IVALL (cuobjdump): a GPU-scope fence per column, then a second
//     DSMEM round trip to fetch the winner's row, then a physical row swap (two more block barriers);
//   * every rank-1 update streams the 256 x 64 slab through shared memory (LDS + STS per DFMA: ~1000 cycles per column).
// This kernel removes all three:
//   * one thread owns one row and keeps it in REGISTERS (64 doubles; the column loop is unrolled 8-wide and the register window is
//     shifted by 8 after each block, so every index is static); finished L columns go to a shared-memory slab once per block;
//   * the exchange is `st.async ... mbarrier::complete_tx::bytes` (SASS: STAS.128): every CTA PUSHES {|candidate| key, row id,
//     candidate row[c..64)} into a mailbox in EVERY peer's shared memory and each peer's mbarrier counts the bytes; a CTA waits
//     (mbarrier.try_wait, hardware sleep) on its own barrier only -- one one-way DSMEM latency per column, no fences, no remote loads;
//   * pivoting is IMPLICIT: rows never move between threads. The winner's thread just stops updating (its registers now hold the
//     U row); the LAPACK swap sequence is replayed on 64-entry position tables after the column loop and only decides WHERE each
//     thread writes its row back, plus the net (dst <- src) move list the other columns are permuted with.
// Pivot choice = max |a(r,c)| over the live rows, ties -> smallest original row: deterministic. (LAPACK breaks ties by current
// position; the two differ only for bit-equal |values|.) The multipliers are a * (1/pivot) instead of a / pivot (<= 1 ulp apart).
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

namespace rm {

constexpr int LU_NB = 64;
// Net row permutation of one panel: row dst[k] receives the row that was at src[k] (global row indices), k < count <= 2*NB.
struct RowMoves {
  uint32_t count;
  unsigned long long dst[2 * LU_NB];
  unsigned long long src[2 * LU_NB];
};

namespace lupush {

constexpr int NB = LU_NB, ROWS = 256, MAXC = 16, BLK = 8;
constexpr int INFO_SINGULAR = 1, INFO_TIMEOUT = 3;

struct __align__(16) Smem {
  double slab[ROWS][NB + 1];               // finished columns of every local row (written once per 8-column block; +1: conflict-free)
  double pad_;                             // keeps mb_row 16-byte aligned (ROWS*(NB+1) is even, so this is only a guard for edits)
  double pad2_;
  double mb_row[2][MAXC][NB];              // [column parity][sender][window-relative column]: candidate rows pushed by every CTA
  unsigned long long mb_meta[2][MAXC][2];  // {key, row}
  double stage[NB];                        // the local candidate's register window, staged by its owner thread for the push
  unsigned long long wkey[2][ROWS / 32];   // per-warp candidates of the next column
  unsigned wrow[2][ROWS / 32];
  unsigned long long bar[2];               // mbarriers, one per column parity
  unsigned long long piv_key[NB];          // replay input: key and original row of every pivot
  unsigned piv_home[NB];
  unsigned cur[NB], inv[NB];               // replay: current position of original row d < NB; original row at position c < NB
  unsigned char was_pivot[NB];
  unsigned fin[ROWS];                      // final panel-local position of every local row
  int abort_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
// 16-byte store into a peer's shared memory that also credits 16 bytes to the peer's mbarrier when it lands
__device__ __forceinline__ void st_async16(uint32_t remote, unsigned long long a, unsigned long long b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote), "l"(a), "l"(b), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void argmax3(unsigned long long& key, unsigned& row) {  // max key, ties -> smallest row; result in every lane
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  row = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? row : 0xffffffffu);
  key = ((unsigned long long)mh << 32) | ml;
}
// order-preserving key of |v| for a live row: 0 = no candidate, 1 = zero or NaN (never preferred over a real value)
__device__ __forceinline__ unsigned long long cand_key(double v, bool live) {
  const double av = fabs(v);
  return !live ? 0ull : (av == av ? (unsigned long long)__double_as_longlong(av) + 1ull : 1ull);
}

// One cluster = the whole grid (gridDim.x = cluster size <= 16, a power of two >= ceil((n - j0) / 256)); exactly NB columns.
// info[0]: INFO_SINGULAR when a pivot is exactly zero, INFO_TIMEOUT when the bounded mailbox wait expired (protocol failure: the
// host turns it into an error; later panels return at once). piv_minmax: running min / max |pivot| for the conditioning gate.
__global__ void __launch_bounds__(ROWS, 1)
lu_panel_push_kernel(double* __restrict__ A, uint64_t lda, uint64_t n, uint64_t j0, unsigned long long* __restrict__ ipiv, int* __restrict__ info,
                     double* __restrict__ piv_minmax, RowMoves* __restrict__ moves) {
  extern __shared__ __align__(16) unsigned char lupush_smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(lupush_smem_raw);
  if (*(volatile int*)info == INFO_TIMEOUT) return;  // set by an earlier panel: uniform over the cluster (stream order)
  const unsigned nblk = gridDim.x, me = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t m = n - j0;
  const unsigned base = me * ROWS, r = base + tid;  // panel-local original row owned by this thread
  const bool valid = r < m;
  double* P = A + j0 + j0 * lda;

  double a[NB];  // register window: a[k] = column (8*b + k) of this thread's row while block b is being factored
#pragma unroll
  for (int e = 0; e < NB; ++e) a[e] = valid ? P[r + (uint64_t)e * lda] : 0.0;
  S.fin[tid] = r;
  if (tid < NB) { S.cur[tid] = tid; S.inv[tid] = tid; S.was_pivot[tid] = 0; }
  if (tid == 0) {
    S.abort_flag = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  bool live = valid;
  {  // per-warp candidates of column 0
    unsigned long long key = cand_key(a[0], live);
    unsigned row = r;
    argmax3(key, row);
    if (lane == 0) { S.wkey[0][warp] = key; S.wrow[0][warp] = row; }
  }
  cooperative_groups::this_cluster().sync();  // barriers initialised everywhere before the first push; also the block barrier for wkey

  for (int b = 0; b < NB / BLK; ++b) {
    const int len = NB - BLK * b;  // live width of the register window
#pragma unroll
    for (int k = 0; k < BLK; ++k) {
      const int c = BLK * b + k;
      const int buf = k & 1;
      const unsigned parity = (unsigned)(k >> 1) & 1u;  // use index of bar[buf] is c >> 1 = 4*b + (k >> 1)
      const int e_lo = k & ~1;                          // first window column that travels (even: 16-byte units)
      // ---- (1) this CTA's candidate: fold the 8 warp candidates (every warp, redundantly) ----
      unsigned long long lkey = lane < ROWS / 32 ? S.wkey[buf][lane] : 0ull;
      unsigned lrow = lane < ROWS / 32 ? S.wrow[buf][lane] : 0xffffffffu;
      argmax3(lkey, lrow);
      const unsigned lp = lkey ? lrow - base : 0u;  // owner thread of the candidate row (thread 0 sends a never-winning dummy otherwise)
      if (tid == 0) {
        const unsigned bytes = nblk * (16u + 8u * (unsigned)(len - e_lo));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.bar[buf])), "r"(bytes) : "memory");
      }
      // ---- (2) the owner's warp pushes {key, row, window[e_lo, len)} into every CTA's mailbox (its own included) ----
      if (warp == (int)(lp >> 5)) {
        if (lane == (int)(lp & 31)) {
#pragma unroll
          for (int e = e_lo; e < NB; e += 2)
            if (e < len) *reinterpret_cast<double2*>(&S.stage[e]) = make_double2(a[e], a[e + 1]);
        }
        __syncwarp();
        const int e0 = 2 * lane;
        if (e0 >= e_lo && e0 < len) {
          const double2 v = *reinterpret_cast<const double2*>(&S.stage[e0]);
          const uint32_t dst = smem_u32(&S.mb_row[buf][me][e0]), bar = smem_u32(&S.bar[buf]);
          for (unsigned pr = 0; pr < nblk; ++pr)
            st_async16(mapa(dst, pr), (unsigned long long)__double_as_longlong(v.x), (unsigned long long)__double_as_longlong(v.y), mapa(bar, pr));
        }
        if ((unsigned)lane < nblk)
          st_async16(mapa(smem_u32(&S.mb_meta[buf][me][0]), lane), lkey, (unsigned long long)(lkey ? lrow : 0xffffffffu), mapa(smem_u32(&S.bar[buf]), lane));
      }
      // ---- (3) wait until every peer's push has landed in OUR shared memory (bounded) ----
      if (!*(volatile int*)&S.abort_flag) {
        const uint32_t bar = smem_u32(&S.bar[buf]);
        const long long t0 = clock64();
        for (;;) {
          uint32_t ok;
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
          if (ok) break;
          if (clock64() - t0 > 200000000LL) { S.abort_flag = 1; atomicExch(info, INFO_TIMEOUT); break; }
        }
      }
      // ---- (4) global pivot (every warp, redundantly) ----
      unsigned long long gkey = (unsigned)lane < nblk ? S.mb_meta[buf][lane][0] : 0ull;
      unsigned grow = (unsigned)lane < nblk ? (unsigned)S.mb_meta[buf][lane][1] : 0xffffffffu;
      argmax3(gkey, grow);
      if (grow >= nblk * ROWS) grow = 0;  // only after a protocol failure; keeps the mailbox index in bounds
      const double* prow = S.mb_row[buf][grow / ROWS];
      if (r == grow) live = false;  // this thread's registers now hold row c of U (and its L part)
      if (tid == 0) { S.piv_key[c] = gkey; S.piv_home[c] = grow; }
      // ---- (5) rank-1 update of the live rows, all in registers ----
      const double pv = prow[k];
      if (live && pv != 0.0) {
        const double l = a[k] * (1.0 / pv);
        a[k] = l;
#pragma unroll
        for (int j = (k + 1) / 2; j < NB / 2; ++j) {
          if (2 * j < len) {
            const double2 u = *reinterpret_cast<const double2*>(&prow[2 * j]);
            if (2 * j > k) a[2 * j] = fma(-l, u.x, a[2 * j]);
            a[2 * j + 1] = fma(-l, u.y, a[2 * j + 1]);
          }
        }
      }
      // ---- (6) block end: park the 8 finished columns, shift the window ----
      if (k == BLK - 1) {
#pragma unroll
        for (int e = 0; e < BLK; ++e) S.slab[tid][BLK * b + e] = a[e];
#pragma unroll
        for (int e = 0; e < NB - BLK; ++e) a[e] = a[e + BLK];
#pragma unroll
        for (int e = NB - BLK; e < NB; ++e) a[e] = 0.0;
      }
      // ---- (7) per-warp candidates of the next column ----
      {
        unsigned long long key = cand_key(a[(k + 1) & (BLK - 1)], live);
        unsigned row = r;
        argmax3(key, row);
        if (lane == 0) { S.wkey[buf ^ 1][warp] = key; S.wrow[buf ^ 1][warp] = row; }
      }
      __syncthreads();
    }
  }

  // ---- replay of the LAPACK swap sequence on position tables (every CTA, one thread: 64 O(1) steps) ----
  if (tid == 0) {
    double pmin = 1.7976931348623157e308, pmax = 0.0;
    bool singular = false;
    for (int c = 0; c < NB; ++c) {
      const unsigned h = S.piv_home[c];          // original row of the pivot
      const unsigned d = S.inv[c];               // original row sitting at position c (always < NB)
      const unsigned p = h < (unsigned)NB ? S.cur[h] : h;  // where the pivot row sits now
      if (p != (unsigned)c) {                    // swap(position c, position p)
        S.cur[d] = p;
        if (p < (unsigned)NB) S.inv[p] = d;
        if (me == 0) S.fin[d] = p;               // original rows < NB live in CTA 0
      }
      if (h < (unsigned)NB) { S.cur[h] = (unsigned)c; S.was_pivot[h] = 1; }
      if (h / ROWS == me) S.fin[h - base] = (unsigned)c;
      if (me == 0) {
        ipiv[j0 + c] = j0 + p;
        const unsigned long long key = S.piv_key[c];
        if (key <= 1ull) singular = true;
        else { const double av = __longlong_as_double((long long)(key - 1ull)); pmin = fmin(pmin, av); pmax = fmax(pmax, av); }
      }
    }
    if (me == 0) {
      if (singular) atomicCAS(info, 0, INFO_SINGULAR);
      if (pmin < piv_minmax[0]) piv_minmax[0] = pmin;
      if (pmax > piv_minmax[1]) piv_minmax[1] = pmax;
      uint32_t cnt = 0;
      for (int c = 0; c < NB; ++c)
        if (S.piv_home[c] != (unsigned)c) { moves->dst[cnt] = j0 + c; moves->src[cnt] = j0 + S.piv_home[c]; ++cnt; }
      for (int d = 0; d < NB; ++d)
        if (!S.was_pivot[d] && S.cur[d] != (unsigned)d) { moves->dst[cnt] = j0 + S.cur[d]; moves->src[cnt] = j0 + d; ++cnt; }
      moves->count = cnt;
    }
  }
  __syncthreads();
  if (valid) {
    const uint64_t fr = S.fin[tid];
#pragma unroll 8
    for (int e = 0; e < NB; ++e) P[fr + (uint64_t)e * lda] = S.slab[tid][e];
  }
  cooperative_groups::this_cluster().sync();  // nobody leaves while a peer's push may still be in flight towards it
}

}  // namespace lupush
}  // namespace rm
