// mc.cu — RNG fills and the fused Monte-Carlo evolution kernel (rows a10/a11).
//
// Reference semantics (CPU baseline): the host evolves the state with `steps` sequential passes, drawing
// `len` normals per pass from ONE global 64-bit LCG (state*6364136223846793005+1, >>11, *2^-53) with
// Box-Muller pairs (element 2j gets r*cos, 2j+1 gets r*sin):
//   crates/runmat-runtime/src/builtins/stats/random/stochastic_evolution.rs:11-32
//   crates/runmat-runtime/src/builtins/common/random.rs:9-13, 271-288, 530-543 (+ jump-ahead :238-257)
// The reference's own GPU kernel (backend/wgpu/shaders/stochastic_evolution.rs:83-117) switches to Philox,
// so its GPU and CPU streams differ by design. Here every thread jumps the SAME LCG to its own position
// (O(log n) affine power, then one precomputed affine hop per step), so the uniforms are bit-identical to the
// host stream and the result is independent of how paths are sharded over GPUs. The whole T-step loop runs
// in registers: S is read once and written once (16 B/path); the kernel is FP64-ALU/SFU bound.
#include "common.h"
#include "mc_math.h"

namespace rm {

namespace {

constexpr uint64_t LCG_MULT = 6364136223846793005ULL;
constexpr uint64_t LCG_INC = 1ULL;

// random.rs:238-257 (advance_state), split into the affine map (mult, plus) so hops can be precomputed.
__host__ __device__ inline void lcg_affine_pow(uint64_t delta, uint64_t* mult, uint64_t* plus) {
  uint64_t cur_mult = LCG_MULT, cur_plus = LCG_INC, acc_mult = 1, acc_plus = 0;
  while (delta > 0) {
    if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = cur_plus * (cur_mult + 1);
    cur_mult = cur_mult * cur_mult;
    delta >>= 1;
  }
  *mult = acc_mult;
  *plus = acc_plus;
}
__host__ __device__ inline uint64_t lcg_advance(uint64_t state, uint64_t delta) {
  uint64_t m, c;
  lcg_affine_pow(delta, &m, &c);
  return m * state + c;
}
__device__ __forceinline__ double lcg_next_uniform(uint64_t& s) {
  s = s * LCG_MULT + LCG_INC;
  return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ uint64_t lcg_next_u53(uint64_t& s) {
  s = s * LCG_MULT + LCG_INC;
  return s >> 11;
}
// One Box-Muller pair as (radius, cos, sin): element 2j gets radius*cos, 2j+1 gets radius*sin (random.rs:271-288).
// LEAN: -2 log(u1) and sincos(2 pi u2) from mc_math.h, computed from the 53-bit LCG integers with a shared-memory table (`tab`,
// filled by load_log_table); false = the CUDA math library on the converted uniforms (RUNMAT_B200_MC_LIBM=1).
template <bool LEAN>
__device__ __forceinline__ void box_muller_polar(uint64_t& s, const double* tab, double& radius, double& cs, double& sn) {
  if (LEAN) {
    const uint64_t x1 = lcg_next_u53(s), x2 = lcg_next_u53(s);
    radius = sqrt(rm_mc::neg2log_u53(x1, tab));
    rm_mc::sincos_turn_u53(x2, &sn, &cs);
  } else {
    double u1 = lcg_next_uniform(s);
    if (u1 <= 0.0) u1 = 2.2250738585072014e-308;  // f64::MIN_POSITIVE
    const double u2 = lcg_next_uniform(s);
    radius = sqrt(-2.0 * log(u1));
    const double angle = 2.0 * 3.14159265358979323846 * u2;
    sincos(angle, &sn, &cs);
  }
}
template <bool LEAN>
__device__ __forceinline__ void box_muller(uint64_t& s, const double* tab, double& z0, double& z1) {
  double radius, cs, sn;
  box_muller_polar<LEAN>(s, tab, radius, cs, sn);
  z0 = radius * cs;
  z1 = radius * sn;
}
// 256 threads copy the 256-entry {1/c, -2 log c} table into shared memory (4 KB): the per-lane lookups are bank-conflict limited
// there, but a divergent index into __constant__ memory would serialise completely.
template <bool LEAN>
__device__ __forceinline__ void load_log_table(double* tab) {
  if (LEAN) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
      *reinterpret_cast<double2*>(&tab[2 * i]) = *reinterpret_cast<const double2*>(&rm_mc::kNeg2LogTab[2 * i]);
    __syncthreads();
  }
}
template <bool LEAN>
__device__ __forceinline__ double mc_exp(double x) { return LEAN ? rm_mc::exp_fast(x) : exp(x); }
inline bool mc_lean() { return getenv("RUNMAT_B200_MC_LIBM") == nullptr; }

template <typename T>
__global__ void uniform_kernel(T* __restrict__ out, uint64_t n, uint64_t state0, uint64_t hop_mult, uint64_t hop_plus) {
  // thread i owns draws i, i+nthr, ...; hop = affine map for nthr steps
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  uint64_t s = lcg_advance(state0, tid + 1);  // state AFTER draw `tid`
  const uint64_t nthr = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < n; i += nthr) {
    out[i] = (T)((double)(s >> 11) * (1.0 / 9007199254740992.0));
    s = hop_mult * s + hop_plus;
  }
}

template <typename T, bool LEAN>
__global__ void normal_kernel(T* __restrict__ out, uint64_t n, uint64_t state0, uint64_t hop_mult, uint64_t hop_plus) {
  __shared__ __align__(16) double tab[512];
  load_log_table<LEAN>(tab);
  const uint64_t pairs = (n + 1) / 2;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= pairs) return;
  uint64_t s = lcg_advance(state0, 2 * tid);  // state BEFORE this pair's first draw
  const uint64_t nthr = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = tid; j < pairs; j += nthr) {
    uint64_t s2 = s;
    double z0, z1;
    box_muller<LEAN>(s2, tab, z0, z1);
    out[2 * j] = (T)z0;
    if (2 * j + 1 < n) out[2 * j + 1] = (T)z1;
    s = hop_mult * s + hop_plus;  // 2*nthr draws ahead
  }
}

// One thread per Box-Muller pair of the GLOBAL path vector. `first_pair` is the global index of this
// launch's first pair; local element e_local = e_global - path_offset.
// SUMLOG: S_T = S_0 * exp(steps*drift + scale * sum_t z_t) with ONE exponential at the end instead of S *= exp(drift + scale z_t)
// every step. Same quantity; the roundings differ by ~sqrt(T) ulp (2e-15 at T = 256, the host's own product carries as much), far
// inside the 1e-10 parity bar, and it removes 2 exps + 2 multiplies + 4 adds per pair-step from an FP64-pipe-bound kernel (r24 ncu:
// math_pipe_throttle is the top stall). Only taken when no partial product can overflow or underflow (the host decides from
// T * (|drift| + 8.6 |scale|) < 600; 8.57 = the largest |z| a 53-bit Box-Muller uniform can produce), so Inf / 0 / NaN patterns of
// the per-step product cannot differ; otherwise, or with RUNMAT_B200_MC_STEPWISE=1, the per-step form runs.
template <typename T, bool LEAN, bool SUMLOG>
__global__ void __launch_bounds__(256, 5)  // 46 registers: the loop keeps its polynomial constants in registers (121 instructions per pair-step instead of 129 at the default 32-register cap)
evolve_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t len, uint64_t path_offset, uint64_t first_pair,
              uint64_t n_pairs, uint64_t state0, uint64_t step_mult, uint64_t step_plus, double drift, double scale, uint32_t steps) {
  __shared__ __align__(16) double tab[512];
  load_log_table<LEAN>(tab);
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_pairs) return;
  const uint64_t pair = first_pair + tid;
  const uint64_t e0 = 2 * pair, e1 = e0 + 1;
  const bool has0 = e0 >= path_offset && e0 - path_offset < len;
  const bool has1 = e1 >= path_offset && e1 - path_offset < len;
  double s0v = has0 ? (double)in[e0 - path_offset] : 0.0;
  double s1v = has1 ? (double)in[e1 - path_offset] : 0.0;
  uint64_t s = lcg_advance(state0, 2 * pair);  // before (u1,u2) of step 0 for this pair
  double a0 = 0.0, a1 = 0.0;
  for (uint32_t t = 0; t < steps; ++t) {
    uint64_t s2 = s;
    if (SUMLOG) {
      // sum_t (drift + scale z_t) = steps*drift + scale * sum_t z_t: the noise sum is one DFMA per path-step
      double radius, cs, sn;
      box_muller_polar<LEAN>(s2, tab, radius, cs, sn);
      a0 = fma(radius, cs, a0);
      a1 = fma(radius, sn, a1);
    } else {
      // stochastic_evolution.rs:25-27: term = drift + scale*noise; value *= exp(term)   (no FMA: -fmad=false)
      double z0, z1;
      box_muller<LEAN>(s2, tab, z0, z1);
      s0v *= mc_exp<LEAN>(drift + scale * z0);
      s1v *= mc_exp<LEAN>(drift + scale * z1);
    }
    s = step_mult * s + step_plus;  // one whole pass (2*ceil(global_len/2) draws) ahead
  }
  if (SUMLOG) {
    const double base = (double)steps * drift;
    s0v *= exp(fma(scale, a0, base));
    s1v *= exp(fma(scale, a1, base));
  }
  if (has0) out[e0 - path_offset] = (T)s0v;
  if (has1) out[e1 - path_offset] = (T)s1v;
}

rm_status evolve(rm_provider* p, const rm_handle* state, double drift, double scale, uint32_t steps, uint64_t path_offset,
                 uint64_t global_len, bool advance_rng, rm_handle* out) {
  void* src;
  uint64_t len;
  RM_TRY(resolve(p, state, &src, &len));
  RM_REQUIRE(path_offset + len <= global_len, RM_INVALID_ARG, "stochastic_evolution: shard [%llu, %llu) exceeds global length %llu",
             (unsigned long long)path_offset, (unsigned long long)(path_offset + len), (unsigned long long)global_len);
  void* dst;
  RM_TRY(alloc_tensor(p, state->shape, state->rank, out, &dst));
  if (len == 0) return RM_OK;
  uint64_t state0;
  {
    std::lock_guard<std::mutex> lk(p->rng_mu);
    state0 = p->rng_state;
  }
  const uint64_t draws_per_step = 2 * ((global_len + 1) / 2);
  if (steps == 0) {
    RM_CUDA(cudaMemcpyAsync(dst, src, len * p->elem_size(), cudaMemcpyDeviceToDevice, p->stream));
    return RM_OK;
  }
  uint64_t sm, sp;
  lcg_affine_pow(draws_per_step, &sm, &sp);
  const uint64_t first_pair = path_offset / 2;
  const uint64_t last_pair = (path_offset + len - 1) / 2;
  const uint64_t n_pairs = last_pair - first_pair + 1;
  const unsigned blocks = (unsigned)((n_pairs + 255) / 256);
  const bool lean = mc_lean();
  const double span = (double)steps * (fabs(drift) + 8.6 * fabs(scale));
  const bool sumlog = lean && span < 600.0 && !getenv("RUNMAT_B200_MC_STEPWISE");  // NaN / Inf parameters fail the comparison -> per-step form
#define RM_EVOLVE(TT, L, SL) evolve_kernel<TT, L, SL><<<blocks, 256, 0, p->stream>>>((const TT*)src, (TT*)dst, len, path_offset, first_pair, n_pairs, state0, sm, sp, drift, scale, steps)
  if (p->precision == RM_F64) { if (sumlog) RM_EVOLVE(double, true, true); else if (lean) RM_EVOLVE(double, true, false); else RM_EVOLVE(double, false, false); }
  else { if (sumlog) RM_EVOLVE(float, true, true); else if (lean) RM_EVOLVE(float, true, false); else RM_EVOLVE(float, false, false); }
#undef RM_EVOLVE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rm_free(p, out); return fail(RM_ERROR, "stochastic_evolution launch failed: %s", cudaGetErrorString(e)); }
  count_launch(p);
  if (advance_rng) {
    std::lock_guard<std::mutex> lk(p->rng_mu);
    p->rng_state = lcg_advance(state0, draws_per_step * (uint64_t)steps);
  }
  return RM_OK;
}

template <bool NORMAL>
rm_status random_fill(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  void* dst;
  RM_TRY(alloc_tensor(p, shape, rank, out, &dst));
  const uint64_t n = shape_elems(shape, rank);
  if (n == 0) return RM_OK;
  uint64_t state0;
  const uint64_t draws = NORMAL ? 2 * ((n + 1) / 2) : n;
  {
    std::lock_guard<std::mutex> lk(p->rng_mu);
    state0 = p->rng_state;
    p->rng_state = lcg_advance(state0, draws);
  }
  const uint64_t work = NORMAL ? (n + 1) / 2 : n;
  const uint64_t blocks = std::min<uint64_t>((work + 255) / 256, (uint64_t)p->prop.multiProcessorCount * 8);
  const uint64_t nthr = blocks * 256;
  uint64_t hm, hp;
  lcg_affine_pow(NORMAL ? 2 * nthr : nthr, &hm, &hp);
  if (NORMAL) {
    if (p->precision == RM_F64) (mc_lean() ? normal_kernel<double, true> : normal_kernel<double, false>)<<<(unsigned)blocks, 256, 0, p->stream>>>((double*)dst, n, state0, hm, hp);
    else (mc_lean() ? normal_kernel<float, true> : normal_kernel<float, false>)<<<(unsigned)blocks, 256, 0, p->stream>>>((float*)dst, n, state0, hm, hp);
  } else {
    if (p->precision == RM_F64) uniform_kernel<double><<<(unsigned)blocks, 256, 0, p->stream>>>((double*)dst, n, state0, hm, hp);
    else uniform_kernel<float><<<(unsigned)blocks, 256, 0, p->stream>>>((float*)dst, n, state0, hm, hp);
  }
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

}  // namespace
}  // namespace rm

using namespace rm;

RM_EXPORT rm_status rm_set_rng_state(rm_provider* p, uint64_t state) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  std::lock_guard<std::mutex> lk(p->rng_mu);
  p->rng_state = state == 0 ? 0x9e3779b97f4a7c15ULL : state;  // simple_provider.rs:3627-3640
  return RM_OK;
}
RM_EXPORT rm_status rm_get_rng_state(rm_provider* p, uint64_t* state) {
  RM_REQUIRE(p && state, RM_INVALID_ARG, "get_rng_state: bad arguments");
  std::lock_guard<std::mutex> lk(p->rng_mu);
  *state = p->rng_state;
  return RM_OK;
}
RM_EXPORT rm_status rm_random_uniform(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "random_uniform: bad arguments");
  DeviceGuard g(p->ordinal);
  return random_fill<false>(p, shape, rank, out);
}
RM_EXPORT rm_status rm_random_normal(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "random_normal: bad arguments");
  DeviceGuard g(p->ordinal);
  return random_fill<true>(p, shape, rank, out);
}
RM_EXPORT rm_status rm_stochastic_evolution(rm_provider* p, const rm_handle* state, double drift, double scale, uint32_t steps, rm_handle* out) {
  RM_REQUIRE(p && state && out, RM_INVALID_ARG, "stochastic_evolution: bad arguments");
  DeviceGuard g(p->ordinal);
  return evolve(p, state, drift, scale, steps, 0, handle_elems(state), true, out);
}
RM_EXPORT rm_status rm_stochastic_evolution_sharded(rm_provider* p, const rm_handle* state, double drift, double scale, uint32_t steps,
                                                    uint64_t path_offset, uint64_t global_len, rm_handle* out) {
  RM_REQUIRE(p && state && out, RM_INVALID_ARG, "stochastic_evolution_sharded: bad arguments");
  DeviceGuard g(p->ordinal);
  return evolve(p, state, drift, scale, steps, path_offset, global_len, true, out);
}
RM_EXPORT rm_status rm_payoff_partial_sum(rm_provider* p, const rm_handle* state, double strike, rm_handle* out) {
  RM_REQUIRE(p && state && out, RM_INVALID_ARG, "payoff_partial_sum: bad arguments");
  DeviceGuard g(p->ordinal);
  // sum(max(S - K, 0)) as a one-pass fused reduction (runmat_lcg.m:48: payoff = max(S - K, 0))
  ReductionProgram prog;
  prog.scalar_ty = p->precision == RM_F64 ? "f64" : "f32";
  prog.n_inputs = 1;
  // the strike rides in as the kernel's free scalar p0: one cached kernel serves every strike (and non-finite strikes)
  prog.val_expr = "fmax((v0 - (T)p0), (T)0)";
  uint64_t one[2] = {1, 1};
  const uint64_t n = handle_elems(state);
  if (n == 0) return rm_fill(p, one, 2, 0.0, out);
  return run_reduction_program(p, prog, "payoff", RedOp::Sum, RedLayout::Contig, state, 1, one, 2, n, 1, 1, 0, 1.0, out, nullptr, strike);
}
