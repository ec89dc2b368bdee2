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
//   * the 64 columns are factored in 8-column blocks with DELAYED updates: inside a block a column step only exchanges and updates
//     the block's own <= 8 columns (a 64-byte row per CTA); the winner of each column pushes its whole (stale) register window once,
//     off the critical path, and at the block end every thread folds the 8 pivot rows into its trailing columns with a rank-8
//     update whose multipliers are corrected by the block's 8 x 8 unit-lower factor (l' L = l), so the stale rows need no fix-up;
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
  double pad_[2];                          // keeps the mailboxes 16-byte aligned
  double mb_row[2][MAXC][BLK];             // [column parity][sender]: in-block window of every CTA's candidate row
  unsigned long long mb_meta[2][MAXC][4];  // {key, row, bits of 1/candidate, -}
  double tr_row[2][BLK][NB];               // [block parity][pivot k of the block]: the winner's whole register window at the time it won
  double stage[BLK];                       // the local candidate's in-block window, staged by its owner thread for the push
  double tstage[NB];                       // a winner's whole window, staged for the deferred push
  unsigned long long wkey[2][ROWS / 32];   // per-warp candidates of the next column
  unsigned long long wrcp[2][ROWS / 32];
  unsigned wrow[2][ROWS / 32];
  unsigned long long bar[2];               // mbarriers: in-block exchange, one per column parity
  unsigned long long tbar[2];              //            winners' windows, one per block parity
  unsigned long long piv_key[NB];          // replay input: key and original row of every pivot
  unsigned piv_home[NB];
  unsigned cur[NB], inv[NB];               // replay tables: current position of original row d < NB; original row at position c < NB
  unsigned was_pivot[NB];
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
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait on one of OUR barriers; a protocol failure raises info and makes every later wait of this CTA fall through
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, int* abort_flag, int* info) {
  if (*(volatile int*)abort_flag) return;
  const uint32_t b = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 200000000LL) { *(volatile int*)abort_flag = 1; atomicExch(info, INFO_TIMEOUT); return; }
  }
}
__device__ __forceinline__ unsigned long long max_u64(unsigned long long key) {  // warp-wide, result in every lane
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}
// Lane holding the maximum key (ties -> lowest lane = smallest row: lanes are always ordered by row); same result in every lane.
// One CREDUX on the top word + a ballot; the low word is only consulted when several lanes share the top word (warp-uniform branch).
__device__ __forceinline__ int argmax_lane(unsigned long long key) {
  const unsigned hi = (unsigned)(key >> 32);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  unsigned cand = __ballot_sync(0xffffffffu, hi == mh);
  if (cand & (cand - 1u)) {
    const unsigned lo = (unsigned)key;
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    cand = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
  }
  return __ffs((int)cand) - 1;
}
// 1/v without the library's special-case subroutine: MUFU seed (20 bits) + two Newton steps. Zero / denormal pivots are never
// divided by (the update is skipped for a zero pivot; a denormal pivot fails the host's conditioning gate).
__device__ __forceinline__ double fast_rcp(double v) {
  double y;
  asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));  // volatile: keep it out of the winner-only branch (it overlaps the vote)
  double e = fma(-v, y, 1.0);
  y = fma(y, e, y);
  e = fma(-v, y, 1.0);
  return fma(y, e, y);
}
// order-preserving key of |v| for a live row: 0 = no candidate, 1 = zero or NaN (never preferred over a real value)
__device__ __forceinline__ unsigned long long cand_key(double v, bool live) {
  const double av = fabs(v);
  return !live ? 0ull : (av == av ? (unsigned long long)__double_as_longlong(av) + 1ull : 1ull);
}

// One step of the LAPACK swap sequence on the position tables (O(1), one thread per CTA): pivot c came from original row h.
// Original rows < NB are the only ones that can be displaced without being pivots, so 64-entry tables suffice.
__device__ __forceinline__ void replay_step(Smem& S, int c, unsigned me, unsigned base, uint64_t j0, unsigned long long* __restrict__ ipiv) {
  const unsigned h = S.piv_home[c];                       // original row of pivot c
  const unsigned d = S.inv[c];                            // original row sitting at position c (always < NB)
  const unsigned p = h < (unsigned)NB ? S.cur[h] : h;     // where the pivot row sits now
  if (p != (unsigned)c) {                                 // swap(position c, position p)
    S.cur[d] = p;
    if (p < (unsigned)NB) S.inv[p] = d;
    if (me == 0) S.fin[d] = p;                            // original rows < NB live in CTA 0
  }
  if (h < (unsigned)NB) { S.cur[h] = (unsigned)c; S.was_pivot[h] = 1u; }
  if (h / ROWS == me) S.fin[h - base] = (unsigned)c;
  if (me == 0) ipiv[j0 + c] = j0 + p;
}

// One cluster = the whole grid (gridDim.x = cluster size <= 16, a power of two >= ceil((n - j0) / 256)); exactly NB columns.
// info[0]: INFO_SINGULAR when a pivot is exactly zero, INFO_TIMEOUT when a bounded mailbox wait expired (protocol failure: the
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
  if (tid < NB) { S.cur[tid] = tid; S.inv[tid] = tid; S.was_pivot[tid] = 0u; }
  if (tid == 0) {
    S.abort_flag = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar[1])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.tbar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.tbar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int died = valid ? NB : -1;  // column at which this row became a pivot row (NB: still live, -1: no row)
  int tpush = -1;              // warp-uniform: block-relative column this warp's thread `tlane` won and has not yet pushed its window for
  int tlane = 0;
  {  // per-warp candidates of column 0
    const double rc = fast_rcp(a[0]);
    const unsigned long long key = cand_key(a[0], valid);
    if (lane == argmax_lane(key)) { S.wkey[0][warp] = key; S.wrow[0][warp] = key ? r : 0xffffffffu; S.wrcp[0][warp] = (unsigned long long)__double_as_longlong(rc); }
  }
  cooperative_groups::this_cluster().sync();  // barriers initialised everywhere before the first push; also the block barrier for wkey

  for (int b = 0; b < NB / BLK; ++b) {
    const int len = NB - BLK * b;  // live width of the register window
    const int bp = b & 1;
    const unsigned tparity = (unsigned)(b >> 1) & 1u;
    const bool trailing = len > BLK;
    if (tid == 0 && trailing) mbar_expect(&S.tbar[bp], (unsigned)(BLK * 8 * len));  // 8 winners x their whole window
#pragma unroll
    for (int k = 0; k < BLK; ++k) {
      const int c = BLK * b + k;
      const int buf = k & 1;
      const unsigned parity = (unsigned)(k >> 1) & 1u;  // use index of bar[buf] is c >> 1 = 4*b + (k >> 1)
      const int e_lo = k & ~1;                          // first in-block column that travels (even: 16-byte units)
      const int npairs = (BLK - e_lo) / 2;
      // ---- (1) this CTA's candidate: fold the 8 warp candidates (every warp, redundantly) ----
      const int ow = argmax_lane(lane < ROWS / 32 ? S.wkey[buf][lane] : 0ull);  // the warp that owns the CTA's candidate (warp 0 sends a never-winning dummy when there is none)
      if (tid == 0) mbar_expect(&S.bar[buf], nblk * (32u + 16u * (unsigned)npairs));
      // ---- (2) the owner's warp pushes {key, row, 1/candidate, window[e_lo, 8)} into every CTA's mailbox (its own included) ----
      if (warp == ow) {
        const unsigned long long lkey = S.wkey[buf][ow];
        const unsigned lrow = S.wrow[buf][ow];
        const int ol = lkey ? (int)((lrow - base) & 31u) : 0;
        if (lane == ol) {
#pragma unroll
          for (int e = e_lo; e < BLK; e += 2) *reinterpret_cast<double2*>(&S.stage[e]) = make_double2(a[e], a[e + 1]);
        }
        __syncwarp();
        const unsigned pr = (unsigned)lane & 15u;
        if (pr < nblk) {
          const uint32_t rbar = mapa(smem_u32(&S.bar[buf]), pr);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int pair = (lane >> 4) + 2 * q;  // 0 .. 3
            if (2 * pair >= e_lo) {
              const double2 v = *reinterpret_cast<const double2*>(&S.stage[2 * pair]);
              st_async16(mapa(smem_u32(&S.mb_row[buf][me][2 * pair]), pr), (unsigned long long)__double_as_longlong(v.x),
                         (unsigned long long)__double_as_longlong(v.y), rbar);
            }
          }
          if (lane < 16) st_async16(mapa(smem_u32(&S.mb_meta[buf][me][0]), pr), lkey, (unsigned long long)lrow, rbar);
          else st_async16(mapa(smem_u32(&S.mb_meta[buf][me][2]), pr), S.wrcp[buf][ow], 0ull, rbar);
        }
      }
      // ---- (2b) deferred: the warp that holds the previous column's winner ships that row's whole window (needed at the block end) ----
      if (tpush >= 0) {
        if (lane == tlane) {
#pragma unroll
          for (int e = 0; e < NB; e += 2)
            if (e < len) *reinterpret_cast<double2*>(&S.tstage[e]) = make_double2(a[e], a[e + 1]);
        }
        __syncwarp();
        if (2 * lane < len) {
          const double2 v = *reinterpret_cast<const double2*>(&S.tstage[2 * lane]);
          const uint32_t dst = smem_u32(&S.tr_row[bp][tpush][2 * lane]), tb = smem_u32(&S.tbar[bp]);
          for (unsigned pr = 0; pr < nblk; ++pr)
            st_async16(mapa(dst, pr), (unsigned long long)__double_as_longlong(v.x), (unsigned long long)__double_as_longlong(v.y), mapa(tb, pr));
        }
        tpush = -1;
      }
      // ---- (2c) one thread replays the PREVIOUS column's swap on the position tables while its warp would wait for the mailbox anyway ----
      if (tid == ROWS - 32 && c > 0) replay_step(S, c - 1, me, base, j0, ipiv);
      // ---- (3) wait until every peer's push has landed in OUR shared memory (bounded) ----
      mbar_wait(&S.bar[buf], parity, &S.abort_flag, info);
      // ---- (4) global pivot (every warp, redundantly) ----
      unsigned gb = (unsigned)argmax_lane((unsigned)lane < nblk ? S.mb_meta[buf][lane][0] : 0ull);  // CTA of the winner
      if (gb >= nblk) gb = 0;  // only after a protocol failure
      const unsigned long long gkey = S.mb_meta[buf][gb][0];
      unsigned grow = (unsigned)S.mb_meta[buf][gb][1];
      if (grow >= nblk * ROWS) grow = 0;  // only after a protocol failure; keeps the indices below in bounds
      const double* prow = S.mb_row[buf][gb];
      const double rp = __longlong_as_double((long long)S.mb_meta[buf][gb][2]);
      if (r == grow) died = c;  // this thread's registers now hold row c of U (and its L part)
      if (trailing && gb == me && warp == (int)((grow - base) >> 5)) { tpush = k; tlane = (int)((grow - base) & 31); }
      if (tid == 0) { S.piv_key[c] = gkey; S.piv_home[c] = grow; }
      const bool live = died == NB;
      // ---- (5) rank-1 update of the live rows inside the block, all in registers ----
      if (live && gkey > 1ull) {
        const double l = a[k] * rp;
        a[k] = l;
#pragma unroll
        for (int j = (k + 1) / 2; j < BLK / 2; ++j) {
          const double2 u = *reinterpret_cast<const double2*>(&prow[2 * j]);
          if (2 * j > k) a[2 * j] = fma(-l, u.x, a[2 * j]);
          a[2 * j + 1] = fma(-l, u.y, a[2 * j + 1]);
        }
      }
      // ---- (6) block end: fold the block's 8 pivot rows into the trailing columns, park the finished columns, shift the window ----
      if (k == BLK - 1) {
        if (trailing) {
          if (tpush >= 0) {  // the last winner cannot defer
            if (lane == tlane) {
#pragma unroll
              for (int e = 0; e < NB; e += 2)
                if (e < len) *reinterpret_cast<double2*>(&S.tstage[e]) = make_double2(a[e], a[e + 1]);
            }
            __syncwarp();
            if (2 * lane < len) {
              const double2 v = *reinterpret_cast<const double2*>(&S.tstage[2 * lane]);
              const uint32_t dst = smem_u32(&S.tr_row[bp][tpush][2 * lane]), tb = smem_u32(&S.tbar[bp]);
              for (unsigned pr = 0; pr < nblk; ++pr)
                st_async16(mapa(dst, pr), (unsigned long long)__double_as_longlong(v.x), (unsigned long long)__double_as_longlong(v.y), mapa(tb, pr));
            }
            tpush = -1;
          }
          mbar_wait(&S.tbar[bp], tparity, &S.abort_flag, info);
        }
#pragma unroll
        for (int e = 0; e < BLK; ++e) S.slab[tid][BLK * b + e] = a[e];
        if (trailing && died >= BLK * b) {
          // multipliers of this row against the block's pivots (a row that became pivot j of the block only has j of them), corrected
          // for the fact that the pivot rows in tr_row are STALE (no update of this block applied): l' L = l, L = the block's
          // unit-lower factor = the winners' own in-block multipliers tr_row[.][i][j], i > j.
          const int nmul = died - BLK * b;  // >= 8 for a live row
#pragma unroll
          for (int e = 0; e < BLK; ++e) if (e >= nmul) a[e] = 0.0;
#pragma unroll
          for (int j = BLK - 2; j >= 0; --j) {
#pragma unroll
            for (int i = j + 1; i < BLK; ++i) a[j] = fma(-a[i], S.tr_row[bp][i][j], a[j]);
          }
#pragma unroll
          for (int t = BLK / 2; t < NB / 2; ++t) {
            if (2 * t < len) {
#pragma unroll
              for (int j = 0; j < BLK; ++j) {
                const double2 u = *reinterpret_cast<const double2*>(&S.tr_row[bp][j][2 * t]);
                a[2 * t] = fma(-a[j], u.x, a[2 * t]);
                a[2 * t + 1] = fma(-a[j], u.y, a[2 * t + 1]);
              }
            }
          }
        }
#pragma unroll
        for (int e = 0; e < NB - BLK; ++e) a[e] = a[e + BLK];
#pragma unroll
        for (int e = NB - BLK; e < NB; ++e) a[e] = 0.0;
      }
      // ---- (7) per-warp candidates of the next column (with the reciprocal its winner will be divided by) ----
      {
        const double v = a[(k + 1) & (BLK - 1)];
        const double rc = fast_rcp(v);
        const unsigned long long key = cand_key(v, live);
        if (lane == argmax_lane(key)) { S.wkey[buf ^ 1][warp] = key; S.wrow[buf ^ 1][warp] = key ? r : 0xffffffffu; S.wrcp[buf ^ 1][warp] = (unsigned long long)__double_as_longlong(rc); }
      }
      __syncthreads();
    }
  }

  // ---- the last column's swap, then every row goes home ----
  if (tid == ROWS - 32) replay_step(S, NB - 1, me, base, j0, ipiv);
  __syncthreads();
  if (me == 0 && warp == 0) {
    // net move list (lane l: entries l and l + 32): position c <- pivot c's original row; displaced rows d < NB that never became
    // pivots end at cur[d]. Offsets by ballot prefix sums.
    const unsigned ph0 = S.piv_home[lane], ph1 = S.piv_home[lane + 32];
    const unsigned cur0 = S.cur[lane], cur1 = S.cur[lane + 32];
    const bool a0 = ph0 != (unsigned)lane, a1 = ph1 != (unsigned)(lane + 32);
    const bool b0 = !S.was_pivot[lane] && cur0 != (unsigned)lane, b1 = !S.was_pivot[lane + 32] && cur1 != (unsigned)(lane + 32);
    const unsigned ma0 = __ballot_sync(0xffffffffu, a0), ma1 = __ballot_sync(0xffffffffu, a1);
    const unsigned mb0 = __ballot_sync(0xffffffffu, b0), mb1 = __ballot_sync(0xffffffffu, b1);
    const unsigned lt = (1u << lane) - 1u;
    unsigned off = __popc(ma0 & lt);
    if (a0) { moves->dst[off] = j0 + lane; moves->src[off] = j0 + ph0; }
    off = __popc(ma0) + __popc(ma1 & lt);
    if (a1) { moves->dst[off] = j0 + lane + 32; moves->src[off] = j0 + ph1; }
    off = __popc(ma0) + __popc(ma1) + __popc(mb0 & lt);
    if (b0) { moves->dst[off] = j0 + cur0; moves->src[off] = j0 + lane; }
    off = __popc(ma0) + __popc(ma1) + __popc(mb0) + __popc(mb1 & lt);
    if (b1) { moves->dst[off] = j0 + cur1; moves->src[off] = j0 + lane + 32; }
    if (lane == 0) moves->count = __popc(ma0) + __popc(ma1) + __popc(mb0) + __popc(mb1);
    // extreme |pivot| of the panel from the order-preserving keys
    const unsigned long long k0 = S.piv_key[lane], k1 = S.piv_key[lane + 32];
    const unsigned long long kmax = max_u64(k0 > k1 ? k0 : k1);
    const unsigned long long kmin = ~max_u64(~(k0 < k1 ? k0 : k1));
    if (lane == 0) {
      if (kmin <= 1ull) atomicCAS(info, 0, INFO_SINGULAR);
      else {
        const double pmin = __longlong_as_double((long long)(kmin - 1ull)), pmax = __longlong_as_double((long long)(kmax - 1ull));
        if (pmin < piv_minmax[0]) piv_minmax[0] = pmin;
        if (pmax > piv_minmax[1]) piv_minmax[1] = pmax;
      }
    }
  }
  if (valid) {
    const uint64_t fr = S.fin[tid];
#pragma unroll 8
    for (int e = 0; e < NB; ++e) P[fr + (uint64_t)e * lda] = S.slab[tid][e];
  }
  cooperative_groups::this_cluster().sync();  // nobody leaves while a peer's push may still be in flight towards it
}

}  // namespace lupush
}  // namespace rm
