// gemm.cu — dense column-major matmul C = A*B (+ fused MatmulEpilogue) for the provider.
//
// Replaces the wgpu provider's 32x32 shared-tile, one-output-per-thread WGSL matmul
// (backend/wgpu/shaders/matmul.rs:1-77, provider/ops/linalg/matmul.rs:813-) and the host provider's naive
// j-i-k loop (simple_provider.rs:7698-7741 == builtins/common/linalg.rs:6-32).
//
// Engine 1 (this file): native FP64 tensor-core path. tcgen05.mma has no f64 kind (f16/bf16/tf32/f8f6f4/i8/
// mxf*), so true-f64 products run on the DMMA pipe via mma.sync.m8n8k4.f64. Blocking: 128x128x16 CTA tile,
// 8 warps as 4(M) x 2(N) with 32x64 warp tiles (32 independent DMMA accumulators per k4 step), operands
// staged global->shared with a 4-stage cp.async ring (16-byte copies, zero-filled at the edges), shared
// layouts padded (+4 doubles) so every fragment load is bank-conflict free.
// The epilogue (alpha/beta, row/col scale, clamps, pow, diag capture; order fixed by
// simple_provider.rs:7805-7838) is applied in registers before the single store of C.
// f32 storage (precision F32) uses a plain FFMA register-tiled kernel (not on any benchmark config).
#include "common.h"

namespace rm {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDA_S = BM + 4;  // doubles per k-row of the A tile  (132 % 16 == 4 -> conflict-free)
constexpr int LDB_S = BK + 4;  // doubles per n-row of the B tile  (20 % 16 == 4 -> conflict-free)
constexpr int A_STAGE = BK * LDA_S;
constexpr int B_STAGE = BN * LDB_S;
constexpr size_t GEMM_SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
// TRANSA (C = A' * B, the Gram / SYRK form): the left operand is read k-contiguous like B, so its tile is staged [m][k] like B's
constexpr int A_STAGE_T = BM * LDB_S;
constexpr size_t GEMM_SMEM_T = (size_t)STAGES * (A_STAGE_T + B_STAGE) * sizeof(double);

struct Epilogue {
  double alpha, beta;
  const double* row_scale;
  const double* col_scale;
  int row_div, col_div;
  int has_min, has_max, has_pow;
  double cmin, cmax, pw;
  double* diag;
  int active;  // 0 => plain store
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double apply_epilogue(double v, const Epilogue& ep, uint64_t i, uint64_t j) {
  v = v * ep.alpha + ep.beta;  // -fmad=false: two roundings, as on the host
  if (ep.row_scale) { const double s = ep.row_scale[i]; v = ep.row_div ? v / s : v * s; }
  if (ep.col_scale) { const double s = ep.col_scale[j]; v = ep.col_div ? v / s : v * s; }
  if (ep.has_min) v = fmax(v, ep.cmin);
  if (ep.has_max) v = fmin(v, ep.cmax);
  if (ep.has_pow) v = pow(v, ep.pw);
  if (ep.diag && i == j) ep.diag[i] = v;
  return v;
}

// ALIGNED2: m and k are even -> 16-byte cp.async on both operands; otherwise 8-byte copies.
// TRANSA: A points at a [k x m] column-major matrix (leading dimension lda) and the product is A' * B (no materialised transpose).
template <bool ALIGNED2, bool TRANSA = false>
__global__ void __launch_bounds__(256, 1)
dgemm_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                  uint64_t m, uint64_t n, uint64_t k, uint64_t lda, uint64_t ldb, uint64_t ldc, int subtract,
                  const __grid_constant__ Epilogue ep, const __grid_constant__ OzGuard guard) {
  extern __shared__ __align__(16) double smem[];
  // conditional mode (behind the tcgen05 engine): only tiles its accuracy guard / non-finite scan handed over are computed
  if (guard.flags && !(guard.flags[0] | guard.tileflags[blockIdx.x + (blockIdx.y >> 1) * guard.tiles_m])) return;
  constexpr int ASTG = TRANSA ? A_STAGE_T : A_STAGE;
  double* As = smem;                              // [STAGES][BK][LDA_S]   (TRANSA: [STAGES][BM][LDB_S])
  double* Bs = smem + (size_t)STAGES * ASTG;      // [STAGES][BN][LDB_S]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 3) * 32;   // warp row offset in the CTA tile
  const int wn = (warp >> 2) * 64;  // warp col offset
  // tile rasterisation: x walks M fastest so a wave shares a few B column panels through L2
  const uint64_t m0 = (uint64_t)blockIdx.x * BM, n0 = (uint64_t)blockIdx.y * BN;
  const uint64_t ktiles = (k + BK - 1) / BK;

  auto load_stage = [&](int stage, uint64_t kt) {
    const uint64_t k0 = kt * BK;
    double* as = As + (size_t)stage * ASTG;
    double* bs = Bs + (size_t)stage * B_STAGE;
    if (TRANSA) {
      // A' tile: BM columns (rows of the product) of BK k-contiguous elements, staged exactly like B's tile
      if (ALIGNED2) {
#pragma unroll
        for (int c = tid; c < BM * (BK / 2); c += 256) {
          const int mc = c / (BK / 2), kr = (c % (BK / 2)) * 2;
          const uint64_t gi = m0 + mc, gk = k0 + kr;
          const bool ok = gi < m && gk < k;
          cp_async16(as + mc * LDB_S + kr, A + (ok ? gi * lda + gk : 0), ok);
        }
      } else {
#pragma unroll 4
        for (int c = tid; c < BM * BK; c += 256) {
          const int mc = c / BK, kr = c % BK;
          const uint64_t gi = m0 + mc, gk = k0 + kr;
          const bool ok = gi < m && gk < k;
          cp_async8(as + mc * LDB_S + kr, A + (ok ? gi * lda + gk : 0), ok);
        }
      }
    }
    if (ALIGNED2) {
      // A tile: BK columns of BM rows; 64 16-byte chunks per column
#pragma unroll
      for (int c = tid; !TRANSA && c < BK * (BM / 2); c += 256) {
        const int kc = c / (BM / 2), mr = (c % (BM / 2)) * 2;
        const uint64_t gi = m0 + mr, gk = k0 + kc;
        const bool ok = gi < m && gk < k;
        cp_async16(as + kc * LDA_S + mr, A + (ok ? gk * lda + gi : 0), ok);
      }
      // B tile: BN columns of BK rows; 8 chunks per column
#pragma unroll
      for (int c = tid; c < BN * (BK / 2); c += 256) {
        const int nc = c / (BK / 2), kr = (c % (BK / 2)) * 2;
        const uint64_t gj = n0 + nc, gk = k0 + kr;
        const bool ok = gj < n && gk < k;
        cp_async16(bs + nc * LDB_S + kr, B + (ok ? gj * ldb + gk : 0), ok);
      }
    } else {
#pragma unroll 4
      for (int c = tid; !TRANSA && c < BK * BM; c += 256) {
        const int kc = c / BM, mr = c % BM;
        const uint64_t gi = m0 + mr, gk = k0 + kc;
        const bool ok = gi < m && gk < k;
        cp_async8(as + kc * LDA_S + mr, A + (ok ? gk * lda + gi : 0), ok);
      }
#pragma unroll 4
      for (int c = tid; c < BN * BK; c += 256) {
        const int nc = c / BK, kr = c % BK;
        const uint64_t gj = n0 + nc, gk = k0 + kr;
        const bool ok = gj < n && gk < k;
        cp_async8(bs + nc * LDB_S + kr, B + (ok ? gj * ldb + gk : 0), ok);
      }
    }
  };

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // prologue: fill STAGES-1 stages
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if ((uint64_t)s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  for (uint64_t kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // refill the stage that was consumed in the previous iteration
    const uint64_t next = kt + STAGES - 1;
    if (next < ktiles) load_stage((int)(next % STAGES), next);
    cp_async_commit();

    const double* as = As + (size_t)(kt % STAGES) * ASTG;
    const double* bs = Bs + (size_t)(kt % STAGES) * B_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = TRANSA ? as[(wm + i * 8 + g) * LDB_S + k4 + t] : as[(k4 + t) * LDA_S + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 8; ++j) bf[j] = bs[(wn + j * 8 + g) * LDB_S + k4 + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue + store. Fragment (i,j): rows m0+wm+i*8+g, cols n0+wn+j*8+2t, +1
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t row = m0 + wm + i * 8 + g;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint64_t col = n0 + wn + j * 8 + 2 * t + e;
        if (col < n) {
          double v = acc[i][j][e];
          if (subtract) v = C[row + col * ldc] - v;  // trailing-matrix update of the blocked LU (solve.cu)
          else if (ep.active) v = apply_epilogue(v, ep, row, col);
          C[row + col * ldc] = v;
        }
      }
    }
  }
}

// f32-storage fallback: 64x64 tile, 4x4 micro-tile per thread, FFMA (no config uses it; kept for the
// F32 provider mode the reference defaults to on wgpu: backend/wgpu/provider/init.rs:145-172).
struct EpilogueF {
  float alpha, beta;
  const float* row_scale;
  const float* col_scale;
  int row_div, col_div, has_min, has_max, has_pow;
  float cmin, cmax, pw;
  float* diag;
  int active;
};
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                    uint64_t m, uint64_t n, uint64_t k, const __grid_constant__ EpilogueF ep) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const uint64_t m0 = (uint64_t)blockIdx.x * 64, n0 = (uint64_t)blockIdx.y * 64;
  float acc[4][4] = {};
  for (uint64_t k0 = 0; k0 < k; k0 += 16) {
    for (int c = threadIdx.x; c < 16 * 64; c += 256) {
      const int kc = c / 64, mr = c % 64;
      const uint64_t gi = m0 + mr, gk = k0 + kc;
      As[kc][mr] = (gi < m && gk < k) ? A[gk * m + gi] : 0.f;
      const int nc = c / 16, kr = c % 16;
      const uint64_t gj = n0 + nc, gk2 = k0 + kr;
      Bs[kr][nc] = (gj < n && gk2 < k) ? B[gj * k + gk2] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t row = m0 + tx * 4 + i, col = n0 + ty * 4 + j;
      if (row < m && col < n) {
        float v = acc[i][j];
        if (ep.active) {
          v = v * ep.alpha + ep.beta;
          if (ep.row_scale) v = ep.row_div ? v / ep.row_scale[row] : v * ep.row_scale[row];
          if (ep.col_scale) v = ep.col_div ? v / ep.col_scale[col] : v * ep.col_scale[col];
          if (ep.has_min) v = fmaxf(v, ep.cmin);
          if (ep.has_max) v = fminf(v, ep.cmax);
          if (ep.has_pow) v = powf(v, ep.pw);
          if (ep.diag && row == col) ep.diag[row] = v;
        }
        C[row + col * m] = v;
      }
    }
}

}  // namespace

rm_status matmul_impl(rm_provider* p, const rm_handle* a, const rm_handle* b, const rm_matmul_epilogue* epd, rm_handle* out) {
  // simple_provider.rs:7705-7712
  RM_REQUIRE(a->rank == 2 && b->rank == 2, RM_ERROR, "matmul: only 2D supported");
  const uint64_t m = a->shape[0], k = a->shape[1], kb = b->shape[0], n = b->shape[1];
  RM_REQUIRE(k == kb, RM_ERROR, "matmul: inner dims must agree");
  void *pa, *pb;
  RM_TRY(resolve(p, a, &pa, nullptr));
  RM_TRY(resolve(p, b, &pb, nullptr));
  void *prow = nullptr, *pcol = nullptr, *pdiag = nullptr;
  bool active = false, used_tcgen05 = false;
  if (epd) {
    active = !(epd->alpha == 1.0 && epd->beta == 0.0 && !epd->row_scale && !epd->col_scale && !epd->has_clamp_min &&
               !epd->has_clamp_max && !epd->has_pow && !epd->diag_output);  // MatmulEpilogue::is_noop (lib.rs:3540-3549)
    uint64_t ne = 0;
    if (epd->row_scale) { RM_TRY(resolve(p, epd->row_scale, &prow, &ne)); RM_REQUIRE(ne >= m, RM_INVALID_ARG, "matmul_epilogue: row scale has %llu elements, need %llu", (unsigned long long)ne, (unsigned long long)m); }
    if (epd->col_scale) { RM_TRY(resolve(p, epd->col_scale, &pcol, &ne)); RM_REQUIRE(ne >= n, RM_INVALID_ARG, "matmul_epilogue: col scale has %llu elements, need %llu", (unsigned long long)ne, (unsigned long long)n); }
    if (epd->diag_output) {
      RM_TRY(resolve(p, epd->diag_output, &pdiag, &ne));
      RM_REQUIRE(ne >= std::min(m, n), RM_ERROR, "matmul_epilogue: diag_output length %llu insufficient for diag size %llu", (unsigned long long)ne,
                 (unsigned long long)std::min(m, n));  // simple_provider.rs:7792-7801
    }
  }
  uint64_t oshape[2] = {m, n};
  void* pc;
  RM_TRY(alloc_tensor(p, oshape, 2, out, &pc));
  if (m * n == 0) return RM_OK;
  rm_status st = RM_OK;
  if (p->precision == RM_F64) {
    // engine selection: 1 = DMMA, 2 = Ozaki/tcgen05, 0 = auto (tcgen05 once the 128x256 tile grid can fill the SMs)
    const uint64_t oz_tiles = ((m + 127) / 128) * ((n + 255) / 256);
    const bool want_ozaki = p->matmul_engine == 2 || (p->matmul_engine == 0 && oz_tiles >= 96 && k >= 512 && !getenv("RUNMAT_B200_DISABLE_OZAKI"));
    OzGuard guard{};
    std::unique_lock<std::mutex> oz_lock(p->oz_mu, std::defer_lock);
    if (want_ozaki) {
      // The tcgen05 engine never blocks the host: what it cannot do (Inf/NaN, entries whose terms hide under the row/column
      // maximum) is marked in device flags, and the DMMA kernel below runs conditionally on them in the same stream.
      oz_lock.lock();
      bool used = false;
      st = ozaki_matmul(p, (const double*)pa, (const double*)pb, (double*)pc, m, n, k, epd, prow, pcol, pdiag, active, &used, &guard);
      if (st != RM_OK) { std::string msg = last_error(); oz_lock.unlock(); rm_free(p, out); set_error("%s", msg.c_str()); return st; }
      if (!used) { guard = OzGuard{}; oz_lock.unlock(); }
      used_tcgen05 = used;
    }
    Epilogue ep{};
    ep.alpha = 1.0;
    if (epd) {
      ep = Epilogue{epd->alpha, epd->beta, (const double*)prow, (const double*)pcol, epd->row_op == RM_SCALE_DIVIDE, epd->col_op == RM_SCALE_DIVIDE,
                    epd->has_clamp_min, epd->has_clamp_max, epd->has_pow, epd->clamp_min, epd->clamp_max, epd->pow_exponent, (double*)pdiag, active ? 1 : 0};
    }
    dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN));
    if (grid.y > 65535) { rm_free(p, out); return fail(RM_UNSUPPORTED, "matmul: n too large for this kernel"); }
    const bool aligned = (m % 2 == 0) && (k % 2 == 0);
    if (aligned) {
      cudaFuncSetAttribute(dgemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
      dgemm_dmma_kernel<true><<<grid, 256, GEMM_SMEM, p->stream>>>((const double*)pa, (const double*)pb, (double*)pc, m, n, k, m, k, m, 0, ep, guard);
    } else {
      cudaFuncSetAttribute(dgemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
      dgemm_dmma_kernel<false><<<grid, 256, GEMM_SMEM, p->stream>>>((const double*)pa, (const double*)pb, (double*)pc, m, n, k, m, k, m, 0, ep, guard);
    }
  } else {
    EpilogueF ep{};
    ep.alpha = 1.f;
    if (epd) {
      ep = EpilogueF{(float)epd->alpha, (float)epd->beta, (const float*)prow, (const float*)pcol, epd->row_op == RM_SCALE_DIVIDE, epd->col_op == RM_SCALE_DIVIDE,
                     epd->has_clamp_min, epd->has_clamp_max, epd->has_pow, (float)epd->clamp_min, (float)epd->clamp_max, (float)epd->pow_exponent, (float*)pdiag, active ? 1 : 0};
    }
    dim3 grid((unsigned)((m + 63) / 64), (unsigned)((n + 63) / 64));
    sgemm_kernel<<<grid, 256, 0, p->stream>>>((const float*)pa, (const float*)pb, (float*)pc, m, n, k, ep);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { st = fail(RM_ERROR, "matmul launch failed: %s", cudaGetErrorString(e)); rm_free(p, out); }
  else {
    count_launch(p);
    // shape keys as record_matmul_kernel_launch logs them (backend/wgpu/provider/helpers.rs:36-50); tuning: this backend's engine
    record_launch(p, "matmul", {{"m", m}, {"n", n}, {"k", k}}, {{"tcgen05", used_tcgen05 ? 1ull : 0ull}, {"epilogue", epd ? 1ull : 0ull}});
  }
  return st;
}

// Skinny-update kernel for the blocked LU (solve.cu): C -= A * B with a 64 x 64 x 16 CTA tile, 8 warps as 2(M) x 4(N) with
// 32 x 16 warp tiles. The LU's in-block, U12 and back-substitution updates are k = 64 products with a 64..256-wide N: on the
// 128 x 128 kernel above they fill 30-60 CTAs, each of which is DMMA-bound for 8+ us (a 128 x 128 x 16 step is 2.1 us of one SM's
// FP64 tensor pipe) and ran at a ~30 us floor (r17 launch list: 181 such launches = 6.4 ms of a 27 ms solve). Quartering the tile
// spreads the same flops over 4x the SMs. Aligned operands only (even leading dimensions, 16-byte pointers): the caller falls
// back to the large kernel otherwise.
constexpr int SBM = 64, SBN = 64;
constexpr int SLDA_S = SBM + 4;
constexpr int SA_STAGE = BK * SLDA_S, SB_STAGE = SBN * LDB_S;
constexpr size_t SGEMM_SMEM = (size_t)STAGES * (SA_STAGE + SB_STAGE) * sizeof(double);
__global__ void __launch_bounds__(256)
dgemm_sub_small_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, uint64_t m, uint64_t n, uint64_t k,
                       uint64_t lda, uint64_t ldb, uint64_t ldc) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                               // [STAGES][BK][SLDA_S]
  double* Bs = smem + (size_t)STAGES * SA_STAGE;   // [STAGES][SBN][LDB_S]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16;
  const uint64_t m0 = (uint64_t)blockIdx.x * SBM, n0 = (uint64_t)blockIdx.y * SBN;
  const uint64_t ktiles = (k + BK - 1) / BK;
  auto load_stage = [&](int stage, uint64_t kt) {
    const uint64_t k0 = kt * BK;
    double* as = As + (size_t)stage * SA_STAGE;
    double* bs = Bs + (size_t)stage * SB_STAGE;
#pragma unroll
    for (int c = tid; c < BK * (SBM / 2); c += 256) {
      const int kc = c / (SBM / 2), mr = (c % (SBM / 2)) * 2;
      const uint64_t gi = m0 + mr, gk = k0 + kc;
      const bool ok = gi < m && gk < k;
      cp_async16(as + kc * SLDA_S + mr, A + (ok ? gk * lda + gi : 0), ok);
    }
#pragma unroll
    for (int c = tid; c < SBN * (BK / 2); c += 256) {
      const int nc = c / (BK / 2), kr = (c % (BK / 2)) * 2;
      const uint64_t gj = n0 + nc, gk = k0 + kr;
      const bool ok = gj < n && gk < k;
      cp_async16(bs + nc * LDB_S + kr, B + (ok ? gj * ldb + gk : 0), ok);
    }
  };
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if ((uint64_t)s < ktiles) load_stage(s, s);
    cp_async_commit();
  }
  for (uint64_t kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const uint64_t next = kt + STAGES - 1;
    if (next < ktiles) load_stage((int)(next % STAGES), next);
    cp_async_commit();
    const double* as = As + (size_t)(kt % STAGES) * SA_STAGE;
    const double* bs = Bs + (size_t)(kt % STAGES) * SB_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double af[4], bf[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = as[(k4 + t) * SLDA_S + wm + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 2; ++j) bf[j] = bs[(wn + j * 8 + g) * LDB_S + k4 + t];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t row = m0 + wm + i * 8 + g;
    if (row >= m) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint64_t col = n0 + wn + j * 8 + 2 * t + e;
        if (col < n) C[row + col * ldc] -= acc[i][j][e];
      }
  }
}

// C (ldc) -= A (lda) * B (ldb) on sub-matrices of column-major storage: the GEMM core of mldivide's blocked LU / solves.
rm_status dgemm_sub_strided(rm_provider* p, const double* A, uint64_t lda, const double* B, uint64_t ldb, double* C, uint64_t ldc,
                            uint64_t m, uint64_t n, uint64_t k, cudaStream_t stream) {
  if (!stream) stream = p->stream;
  if (m == 0 || n == 0 || k == 0) return RM_OK;
  Epilogue ep{};
  ep.alpha = 1.0;
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN));
  RM_REQUIRE(grid.y <= 65535, RM_UNSUPPORTED, "gemm update: n too large");
  const bool aligned = (lda % 2 == 0) && (ldb % 2 == 0) && (((uintptr_t)A) % 16 == 0) && (((uintptr_t)B) % 16 == 0);
  if (aligned && (uint64_t)grid.x * grid.y < 2ull * (uint64_t)p->prop.multiProcessorCount && !getenv("RUNMAT_B200_LU_NO_SMALL_GEMM")) {
    // too few 128 x 128 tiles to fill the SMs: the 64 x 64 kernel
    dim3 sgrid((unsigned)((m + SBM - 1) / SBM), (unsigned)((n + SBN - 1) / SBN));
    cudaFuncSetAttribute(dgemm_sub_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SGEMM_SMEM);
    dgemm_sub_small_kernel<<<sgrid, 256, SGEMM_SMEM, stream>>>(A, B, C, m, n, k, lda, ldb, ldc);
  } else if (aligned) {
    cudaFuncSetAttribute(dgemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    dgemm_dmma_kernel<true><<<grid, 256, GEMM_SMEM, stream>>>(A, B, C, m, n, k, lda, ldb, ldc, 1, ep, OzGuard{});
  } else {
    cudaFuncSetAttribute(dgemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    dgemm_dmma_kernel<false><<<grid, 256, GEMM_SMEM, stream>>>(A, B, C, m, n, k, lda, ldb, ldc, 1, ep, OzGuard{});
  }
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}


// out = a' * a (optionally every entry divided by `divisor`, the covariance normalisation) without materialising a'. f64: the
// TRANSA form of the DMMA kernel reads `a` k-contiguous for both operands (the reference's SYRK shader, backend/wgpu/shaders/syrk.rs,
// also reads one input twice). f32 storage keeps the transpose + matmul composition (not on any benchmark config).
rm_status syrk_impl(rm_provider* p, const rm_handle* a, const double* divisor_vec, rm_handle* out) {
  RM_REQUIRE(a->rank == 2, RM_ERROR, "syrk: matrix input required");
  const uint64_t k = a->shape[0], m = a->shape[1];
  if (p->precision != RM_F64 || k == 0 || m == 0) {
    RM_REQUIRE(!divisor_vec, RM_UNSUPPORTED, "syrk: fused normalisation needs f64 storage");
    rm_handle at;
    RM_TRY(rm_transpose(p, a, &at));
    rm_status st = matmul_impl(p, &at, a, nullptr, out);
    std::string msg = st == RM_OK ? "" : last_error();
    rm_free(p, &at);
    if (st != RM_OK) set_error("%s", msg.c_str());
    return st;
  }
  void* pa;
  RM_TRY(resolve(p, a, &pa, nullptr));
  uint64_t oshape[2] = {m, m};
  void* pc;
  RM_TRY(alloc_tensor(p, oshape, 2, out, &pc));
  Epilogue ep{};
  ep.alpha = 1.0;
  if (divisor_vec) { ep.col_scale = divisor_vec; ep.col_div = 1; ep.active = 1; }
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((m + BN - 1) / BN));
  if (grid.y > 65535) { rm_free(p, out); return fail(RM_UNSUPPORTED, "syrk: matrix too wide for this kernel"); }
  if (k % 2 == 0) {
    cudaFuncSetAttribute(dgemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_T);
    dgemm_dmma_kernel<true, true><<<grid, 256, GEMM_SMEM_T, p->stream>>>((const double*)pa, (const double*)pa, (double*)pc, m, m, k, k, k, m, 0, ep, OzGuard{});
  } else {
    cudaFuncSetAttribute(dgemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_T);
    dgemm_dmma_kernel<false, true><<<grid, 256, GEMM_SMEM_T, p->stream>>>((const double*)pa, (const double*)pa, (double*)pc, m, m, k, k, k, m, 0, ep, OzGuard{});
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rm_free(p, out); return fail(RM_ERROR, "syrk launch failed: %s", cudaGetErrorString(e)); }
  count_launch(p);
  record_launch(p, "syrk", {{"rows", k}, {"cols", m}}, {{"transposed_read", 1ull}, {"normalised", divisor_vec ? 1ull : 0ull}});
  return RM_OK;
}

}  // namespace rm

using namespace rm;

RM_EXPORT rm_status rm_matmul(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out) {
  RM_REQUIRE(p && a && b && out, RM_INVALID_ARG, "matmul: bad arguments");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_matmul);
  return matmul_impl(p, a, b, nullptr, out);
}
RM_EXPORT rm_status rm_matmul_epilogue_apply(rm_provider* p, const rm_handle* a, const rm_handle* b, const rm_matmul_epilogue* ep, rm_handle* out) {
  RM_REQUIRE(p && a && b && ep && out, RM_INVALID_ARG, "matmul_epilogue: bad arguments");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_matmul);
  return matmul_impl(p, a, b, ep, out);
}
RM_EXPORT rm_status rm_syrk(rm_provider* p, const rm_handle* a, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "syrk: bad arguments");
  DeviceGuard g(p->ordinal);
  return syrk_impl(p, a, nullptr, out);
}
RM_EXPORT rm_status rm_set_matmul_engine(rm_provider* p, int engine) {
  RM_REQUIRE(p && engine >= 0 && engine <= 2, RM_INVALID_ARG, "set_matmul_engine: engine must be 0, 1 or 2");
  p->matmul_engine = engine;
  return RM_OK;
}
