// ops.cu — C-ABI entry points for the operator surface: fused elementwise / fused reduction (rows a3, a4),
// the unfused elem_* / unary_* / scalar_* ops (row a5) and the reduce_* family (row a6).
//
// Design: the unfused ops are one-node fused programs. They go through the same lowering, the same
// 256-bit-vectorised kernel scaffolding and the same module cache as planner-generated programs, so every
// operator gets the streaming-bandwidth kernel instead of a second, hand-enumerated op-code switch
// (the reference keeps two: BINARY/UNARY/SCALAR_SHADER op-code switches in backend/wgpu/shaders/elementwise.rs
// next to the runtime-generated fused shaders).
#include <algorithm>
#include <memory>

#include "common.h"

using namespace rm;

namespace {

const char* binary_expr(rm_binary_op op) {
  switch (op) {
    case RM_BIN_ADD: return "(v0 + v1)";
    case RM_BIN_SUB: return "(v0 - v1)";
    case RM_BIN_MUL: return "(v0 * v1)";
    case RM_BIN_DIV: return "(v0 / v1)";
    case RM_BIN_POW: return "pow(v0, v1)";
    case RM_BIN_MAX: return "fmax(v0, v1)";   // max.rs:2323-2343: NaN loses unless both NaN
    case RM_BIN_MIN: return "fmin(v0, v1)";
    case RM_BIN_HYPOT: return "hypot(v0, v1)";
    case RM_BIN_ATAN2: return "atan2(v0, v1)";
    case RM_BIN_MOD: return "rm_mod(v0, v1)";
    case RM_BIN_REM: return "rm_rem(v0, v1)";
    case RM_BIN_GE: return "((v0 >= v1) ? (T)1 : (T)0)";
    case RM_BIN_LE: return "((v0 <= v1) ? (T)1 : (T)0)";
    case RM_BIN_LT: return "((v0 < v1) ? (T)1 : (T)0)";
    case RM_BIN_GT: return "((v0 > v1) ? (T)1 : (T)0)";
    case RM_BIN_EQ: return "((v0 == v1) ? (T)1 : (T)0)";
    case RM_BIN_NE: return "((v0 != v1) ? (T)1 : (T)0)";
    default: return nullptr;
  }
}

const char* unary_expr(rm_unary_op op) {
  switch (op) {
    case RM_UN_SIN: return "sin(v0)";
    case RM_UN_COS: return "cos(v0)";
    case RM_UN_TAN: return "tan(v0)";
    case RM_UN_ASIN: return "asin(v0)";
    case RM_UN_ACOS: return "acos(v0)";
    case RM_UN_ATAN: return "atan(v0)";
    case RM_UN_SINH: return "sinh(v0)";
    case RM_UN_COSH: return "cosh(v0)";
    case RM_UN_TANH: return "tanh(v0)";
    case RM_UN_ASINH: return "asinh(v0)";
    case RM_UN_ACOSH: return "acosh(v0)";
    case RM_UN_ATANH: return "atanh(v0)";
    case RM_UN_EXP: return "exp(v0)";
    case RM_UN_EXPM1: return "expm1(v0)";
    case RM_UN_LOG: return "log(v0)";
    case RM_UN_LOG2: return "log2(v0)";
    case RM_UN_LOG10: return "log10(v0)";
    case RM_UN_LOG1P: return "log1p(v0)";
    case RM_UN_SQRT: return "sqrt(v0)";
    case RM_UN_ABS: return "fabs(v0)";
    case RM_UN_SIGN: return "rm_sign(v0)";
    case RM_UN_FLOOR: return "floor(v0)";
    case RM_UN_CEIL: return "ceil(v0)";
    case RM_UN_ROUND: return "round(v0)";  // half away from zero == f64::round
    case RM_UN_FIX: return "trunc(v0)";
    case RM_UN_NEG: return "(-v0)";
    case RM_UN_POW2: return "exp2(v0)";
    case RM_UN_HEAVISIDE: return "rm_heaviside(v0)";
    case RM_UN_SINGLE: return "((T)(float)v0)";
    case RM_UN_DOUBLE: return "v0";
    case RM_UN_ISNAN: return "(rm_isnan(v0) ? (T)1 : (T)0)";
    case RM_UN_ISINF: return "(rm_isinf(v0) ? (T)1 : (T)0)";
    case RM_UN_ISFINITE: return "(rm_isfinite(v0) ? (T)1 : (T)0)";
    case RM_UN_NAN_TO_ZERO: return "(rm_isnan(v0) ? (T)0 : v0)";
    case RM_UN_NOT_NAN_MASK: return "(rm_isnan(v0) ? (T)0 : (T)1)";
    case RM_UN_ERF: return "erf(v0)";
    case RM_UN_GAMMA: return "tgamma(v0)";
    case RM_UN_GAMMALN: return "lgamma(v0)";
    default: return nullptr;
  }
}

const char* scalar_expr(rm_scalar_op op) {
  switch (op) {
    case RM_SC_ADD: return "(v0 + v1)";
    case RM_SC_SUB: return "(v0 - v1)";
    case RM_SC_MUL: return "(v0 * v1)";
    case RM_SC_DIV: return "(v0 / v1)";
    case RM_SC_RSUB: return "(v1 - v0)";
    case RM_SC_RDIV: return "(v1 / v0)";
    case RM_SC_MAX: return "fmax(v0, v1)";
    case RM_SC_MIN: return "fmin(v0, v1)";
    case RM_SC_POW: return "pow(v0, v1)";
    default: return nullptr;
  }
}

ElementwiseProgram one_node(rm_provider* p, uint32_t n_inputs, const char* expr) {
  ElementwiseProgram prog;
  prog.scalar_ty = p->precision == RM_F64 ? "f64" : "f32";
  prog.n_inputs = n_inputs;
  prog.n_outputs = 1;
  prog.outputs.push_back(expr);
  return prog;
}

// broadcast_shapes (builtins/common/broadcast.rs:8-47): front-pad the shorter shape, per-dim expand.
rm_status broadcast_shape(const rm_handle* a, const rm_handle* b, uint64_t* out, uint32_t* rank) {
  const uint32_t r = std::max(a->rank, b->rank);
  for (uint32_t d = 0; d < r; ++d) {
    const uint64_t x = d < r - a->rank ? 1 : a->shape[d - (r - a->rank)];
    const uint64_t y = d < r - b->rank ? 1 : b->shape[d - (r - b->rank)];
    if (x == y) out[d] = x;
    else if (x == 1) out[d] = y;
    else if (y == 1) out[d] = x;
    else if (x == 0 || y == 0) out[d] = 0;
    else return fail(RM_ERROR, "size mismatch between inputs (dimension %u has lengths %llu and %llu)", d + 1, (unsigned long long)x, (unsigned long long)y);
  }
  *rank = r;
  return RM_OK;
}

rm_status empty_result(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) { return alloc_tensor(p, shape, rank, out, nullptr); }

// Parsed-program cache keyed by a 64-bit hash of the shader text (+length): a repeated fused call costs one hash of
// the text instead of a re-parse, and the kernel-cache key stays short. Mirrors the reference's thread-local plan
// cache (fusion.rs:679-747) on the provider side.
uint64_t text_hash(const char* s, size_t* len_out) {
  uint64_t h = 1469598103934665603ULL;
  size_t n = 0;
  for (; s[n]; ++n) { h ^= (unsigned char)s[n]; h *= 1099511628211ULL; }
  *len_out = n;
  return h;
}
// The cache is keyed by the FULL shader text (a 64-bit hash alone could silently run the wrong kernel on a collision); the
// short key handed to the kernel cache is hash:len, made unique with a suffix if two different texts ever share it.
template <typename Prog>
struct ParsedCache {
  struct Entry { std::shared_ptr<Prog> prog; std::string key; };
  std::mutex mu;
  std::unordered_map<std::string, Entry> map;            // shader text -> parsed program + short key
  std::unordered_map<std::string, uint32_t> short_keys;  // hash:len -> number of distinct texts seen with it
  static constexpr size_t kMaxEntries = 4096;            // bounded: a caller streaming unique shaders cannot grow it forever
};
ParsedCache<ElementwiseProgram> g_ew_cache;
ParsedCache<ReductionProgram> g_red_cache;

template <typename Prog, typename ParseFn>
rm_status parsed(ParsedCache<Prog>& cache, const char* shader, ParseFn parse, std::shared_ptr<Prog>* out, std::string* key) {
  const std::string text(shader);
  {
    std::lock_guard<std::mutex> lk(cache.mu);
    auto it = cache.map.find(text);
    if (it != cache.map.end()) { *out = it->second.prog; *key = it->second.key; return RM_OK; }
  }
  auto prog = std::make_shared<Prog>();
  std::string err;
  if (!parse(shader, prog.get(), &err)) return fail(RM_COMPILE_ERROR, "%s", err.c_str());
  size_t len;
  const uint64_t h = text_hash(shader, &len);
  char buf[64];
  snprintf(buf, sizeof buf, "%016llx:%zu", (unsigned long long)h, len);
  std::lock_guard<std::mutex> lk(cache.mu);
  auto it = cache.map.find(text);  // another thread may have parsed the same text meanwhile
  if (it != cache.map.end()) { *out = it->second.prog; *key = it->second.key; return RM_OK; }
  if (cache.map.size() >= ParsedCache<Prog>::kMaxEntries) cache.map.clear();  // short_keys is kept: suffixes stay unique
  const uint32_t dup = cache.short_keys[buf]++;
  std::string k = buf;
  if (dup) k += "#" + std::to_string(dup);
  cache.map[text] = {prog, k};
  *out = prog;
  *key = k;
  return RM_OK;
}

}  // namespace

// =============================================================================================================
// a5: unfused operator surface
// =============================================================================================================
RM_EXPORT rm_status rm_elem_binary(rm_provider* p, rm_binary_op op, const rm_handle* a, const rm_handle* b, rm_handle* out) {
  RM_REQUIRE(p && a && b && out, RM_INVALID_ARG, "elem_binary: bad arguments");
  const char* expr = binary_expr(op);
  RM_REQUIRE(expr, RM_UNSUPPORTED, "elem_binary: op %d not supported by provider", (int)op);
  DeviceGuard g(p->ordinal);
  uint64_t shape[RM_MAX_RANK];
  uint32_t rank;
  RM_TRY(broadcast_shape(a, b, shape, &rank));
  const uint64_t len = shape_elems(shape, rank);
  if (len == 0) return empty_result(p, shape, rank, out);
  rm_handle in[2] = {*a, *b};
  return run_elementwise_program(p, one_node(p, 2, expr), std::string("bin:") + expr, in, 2, shape, rank, len, out);
}

RM_EXPORT rm_status rm_unary(rm_provider* p, rm_unary_op op, const rm_handle* a, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "unary: bad arguments");
  const char* expr = unary_expr(op);
  RM_REQUIRE(expr, RM_UNSUPPORTED, "unary: op %d not supported by provider", (int)op);
  DeviceGuard g(p->ordinal);
  const uint64_t len = handle_elems(a);
  if (len == 0) return empty_result(p, a->shape, a->rank, out);
  return run_elementwise_program(p, one_node(p, 1, expr), std::string("un:") + expr, a, 1, a->shape, a->rank, len, out);
}

RM_EXPORT rm_status rm_scalar_op_apply(rm_provider* p, rm_scalar_op op, const rm_handle* a, double scalar, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "scalar op: bad arguments");
  const char* expr = scalar_expr(op);
  RM_REQUIRE(expr, RM_UNSUPPORTED, "scalar op %d not supported by provider", (int)op);
  DeviceGuard g(p->ordinal);
  const uint64_t len = handle_elems(a);
  if (len == 0) return empty_result(p, a->shape, a->rank, out);
  // the scalar rides along as a 1-element device tensor, exactly how the fusion executor feeds constants
  // (fusion_exec.rs:305-326); it is hoisted to a register by the Flat kernel variant.
  uint64_t one[2] = {1, 1};
  rm_handle sh;
  RM_TRY(rm_fill(p, one, 2, scalar, &sh));
  rm_handle in[2] = {*a, sh};
  rm_status st = run_elementwise_program(p, one_node(p, 2, expr), std::string("sc:") + expr, in, 2, a->shape, a->rank, len, out);
  std::string msg = st == RM_OK ? "" : last_error();
  rm_free(p, &sh);
  if (st != RM_OK) set_error("%s", msg.c_str());
  return st;
}

#define RM_BIN_WRAPPER(name, op) \
  RM_EXPORT rm_status name(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out) { return rm_elem_binary(p, op, a, b, out); }
RM_BIN_WRAPPER(rm_elem_add, RM_BIN_ADD)
RM_BIN_WRAPPER(rm_elem_mul, RM_BIN_MUL)
RM_BIN_WRAPPER(rm_elem_max, RM_BIN_MAX)
RM_BIN_WRAPPER(rm_elem_min, RM_BIN_MIN)
RM_BIN_WRAPPER(rm_elem_sub, RM_BIN_SUB)
RM_BIN_WRAPPER(rm_elem_div, RM_BIN_DIV)
RM_BIN_WRAPPER(rm_elem_pow, RM_BIN_POW)
RM_BIN_WRAPPER(rm_elem_hypot, RM_BIN_HYPOT)
RM_BIN_WRAPPER(rm_elem_atan2, RM_BIN_ATAN2)
#define RM_UN_WRAPPER(name, op) \
  RM_EXPORT rm_status name(rm_provider* p, const rm_handle* a, rm_handle* out) { return rm_unary(p, op, a, out); }
RM_UN_WRAPPER(rm_unary_sin, RM_UN_SIN)
RM_UN_WRAPPER(rm_unary_cos, RM_UN_COS)
RM_UN_WRAPPER(rm_unary_tan, RM_UN_TAN)
RM_UN_WRAPPER(rm_unary_tanh, RM_UN_TANH)
RM_UN_WRAPPER(rm_unary_exp, RM_UN_EXP)
RM_UN_WRAPPER(rm_unary_log, RM_UN_LOG)
RM_UN_WRAPPER(rm_unary_sqrt, RM_UN_SQRT)
RM_UN_WRAPPER(rm_unary_abs, RM_UN_ABS)
RM_UN_WRAPPER(rm_unary_floor, RM_UN_FLOOR)
RM_UN_WRAPPER(rm_unary_round, RM_UN_ROUND)
#define RM_SC_WRAPPER(name, op) \
  RM_EXPORT rm_status name(rm_provider* p, const rm_handle* a, double s, rm_handle* out) { return rm_scalar_op_apply(p, op, a, s, out); }
RM_SC_WRAPPER(rm_scalar_add, RM_SC_ADD)
RM_SC_WRAPPER(rm_scalar_sub, RM_SC_SUB)
RM_SC_WRAPPER(rm_scalar_mul, RM_SC_MUL)
RM_SC_WRAPPER(rm_scalar_div, RM_SC_DIV)
RM_SC_WRAPPER(rm_scalar_rsub, RM_SC_RSUB)
RM_SC_WRAPPER(rm_scalar_rdiv, RM_SC_RDIV)
RM_SC_WRAPPER(rm_scalar_max, RM_SC_MAX)
RM_SC_WRAPPER(rm_scalar_min, RM_SC_MIN)

// =============================================================================================================
// a3: fused elementwise
// =============================================================================================================
RM_EXPORT rm_status rm_fused_elementwise_multi(rm_provider* p, const char* shader, const rm_handle* inputs, uint32_t n_inputs,
                                               const uint64_t* output_shape, uint32_t rank, uint64_t len, uint32_t num_outputs, rm_handle* outs) {
  RM_REQUIRE(p && shader && outs && (inputs || n_inputs == 0), RM_INVALID_ARG, "fused_elementwise: bad arguments");
  RM_REQUIRE(n_inputs > 0, RM_ERROR, "fused_elementwise: no inputs");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_fused_elementwise);
  std::shared_ptr<ElementwiseProgram> prog;
  std::string key;
  RM_TRY(parsed(g_ew_cache, shader, parse_elementwise_wgsl, &prog, &key));
  RM_REQUIRE(prog->n_outputs == num_outputs, RM_INVALID_ARG, "fused_elementwise: shader writes %u outputs, caller expects %u", prog->n_outputs, num_outputs);
  rm_status st = run_elementwise_program(p, *prog, "wgsl:" + key, inputs, n_inputs, output_shape, rank, len, outs);
  if (st == RM_OK) {
    // names / shape keys as the wgpu provider logs them (backend/wgpu/provider/ops/telemetry.rs:26-61)
    if (num_outputs == 1) record_launch(p, "fused_elementwise", {{"len", len}, {"inputs", n_inputs}, {"rank", rank}}, {{"block", 256}, {"vec_bytes", 32}});
    else record_launch(p, "fused_elementwise_multi", {{"len", len}, {"inputs", n_inputs}, {"rank", rank}, {"num_outputs", num_outputs}}, {{"block", 256}, {"vec_bytes", 32}});
  }
  return st;
}

RM_EXPORT rm_status rm_fused_elementwise(rm_provider* p, const char* shader, const rm_handle* inputs, uint32_t n_inputs,
                                         const uint64_t* output_shape, uint32_t rank, uint64_t len, rm_handle* out) {
  return rm_fused_elementwise_multi(p, shader, inputs, n_inputs, output_shape, rank, len, 1, out);
}

// =============================================================================================================
// a4: fused reduction
// =============================================================================================================
RM_EXPORT rm_status rm_fused_reduction(rm_provider* p, const char* shader, const rm_handle* inputs, uint32_t n_inputs,
                                       const uint64_t* output_shape, uint32_t rank, uint64_t reduce_len, uint64_t num_slices,
                                       uint32_t /*workgroup_size: a wgpu tuning hint; CUDA geometry is chosen here*/,
                                       rm_reduction_flavor flavor, double custom_scale, rm_handle* out) {
  RM_REQUIRE(p && shader && inputs && out, RM_INVALID_ARG, "fused_reduction: bad arguments");
  RM_REQUIRE(reduce_len > 0, RM_ERROR, "fused_reduction: zero reduce_len");  // fusion_exec.rs: the executor rejects it before dispatch
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_fused_reduction);
  std::shared_ptr<ReductionProgram> progp;
  std::string key;
  RM_TRY(parsed(g_red_cache, shader, parse_reduction_wgsl, &progp, &key));
  const ReductionProgram& prog = *progp;
  // ReductionFlavor::scale (lib.rs:876-887)
  int use_div = 0;
  double factor = 1.0;
  if (flavor == RM_FLAVOR_MEAN) { use_div = reduce_len ? 1 : 0; factor = reduce_len ? (double)reduce_len : 1.0; }
  else if (flavor == RM_FLAVOR_CUSTOM) factor = custom_scale;
  const RedLayout layout = prog.axis == 0 ? RedLayout::Contig : RedLayout::Strided;
  rm_status st = run_reduction_program(p, prog, "wgsl:" + key, RedOp::Sum, layout, inputs, n_inputs, output_shape, rank,
                                       reduce_len, num_slices, /*inner=*/num_slices, use_div, factor, out);
  if (st == RM_OK)  // backend/wgpu/provider/ops/telemetry.rs:126-146
    record_launch(p, "fused_reduction", {{"reduce_len", reduce_len}, {"slices", num_slices}, {"rank", rank}}, {{"block", 256}, {"flavor", (uint64_t)flavor}, {"axis", (uint64_t)prog.axis}});
  return st;
}

// Sharded form: every rank reduces its own inputs to ONE scalar and the scalars are summed over the ranks of the
// peer-memory exchange (rm_comm_p2p_connect). The publish into the peers' slots is the tail of the reduction kernel's last
// block; `out` (1x1) holds the global sum once the lazy combine has run (waited for at first use, like an upload).
RM_EXPORT rm_status rm_fused_reduction_allreduce(rm_provider* p, const char* shader, const rm_handle* inputs, uint32_t n_inputs, uint64_t reduce_len,
                                                 rm_reduction_flavor flavor, double custom_scale, rm_handle* out) {
  RM_REQUIRE(p && shader && inputs && out, RM_INVALID_ARG, "fused_reduction_allreduce: bad arguments");
  RM_REQUIRE(reduce_len > 0, RM_ERROR, "fused_reduction: zero reduce_len");
  RM_REQUIRE(p->precision == RM_F64, RM_UNSUPPORTED, "fused_reduction_allreduce: f64 providers only");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_fused_reduction);
  std::shared_ptr<ReductionProgram> progp;
  std::string key;
  RM_TRY(parsed(g_red_cache, shader, parse_reduction_wgsl, &progp, &key));
  int use_div = 0;
  double factor = 1.0;
  if (flavor == RM_FLAVOR_MEAN) { use_div = 1; factor = (double)reduce_len; }
  else if (flavor == RM_FLAVOR_CUSTOM) factor = custom_scale;
  std::lock_guard<std::mutex> lk(p->comm_mu);
  P2PPublish pub;
  RM_REQUIRE(p2p_begin(p, &pub), RM_ERROR, "fused_reduction_allreduce: rm_comm_p2p_connect has not been called");
  uint64_t one[2] = {1, 1};
  rm_handle local;
  RM_TRY(run_reduction_program(p, *progp, "wgsl:" + key, RedOp::Sum, RedLayout::Contig, inputs, n_inputs, one, 2, reduce_len, 1, 1, use_div, factor, &local, &pub));
  rm_status st = p2p_finish(p, out);
  std::string msg = st == RM_OK ? "" : last_error();
  rm_free(p, &local);
  if (st != RM_OK) set_error("%s", msg.c_str());
  return st;
}

// =============================================================================================================
// a6: reductions
// =============================================================================================================
namespace {

ReductionProgram red_program(rm_provider* p, const char* val, bool omit_nan) {
  ReductionProgram r;
  r.scalar_ty = p->precision == RM_F64 ? "f64" : "f32";
  r.n_inputs = 1;
  r.val_expr = val;
  r.omit_nan = omit_nan;
  return r;
}

// Reduce the contiguous dim range [d0, d1] (zero-based, inclusive) of `a`.
// pre = prod(shape[:d0]), n = prod(shape[d0..d1]), post = prod(shape[d1+1:]).
rm_status reduce_range(rm_provider* p, const rm_handle* a, uint32_t d0, uint32_t d1, RedOp op, const char* val, int mean, rm_handle* out) {
  uint64_t pre = 1, n = 1, post = 1;
  uint64_t oshape[RM_MAX_RANK];
  const uint32_t rank = std::max<uint32_t>(a->rank, 2);
  for (uint32_t d = 0; d < rank; ++d) {
    const uint64_t e = d < a->rank ? a->shape[d] : 1;
    if (d < d0) pre *= e;
    else if (d <= d1) n *= e;
    else post *= e;
    oshape[d] = (d >= d0 && d <= d1) ? 1 : e;
  }
  const uint64_t slices = pre * post;
  if (slices == 0) return alloc_tensor(p, oshape, rank, out, nullptr);
  RM_REQUIRE(n > 0, RM_UNSUPPORTED, "reduction over an empty dimension not supported by provider");
  ReductionProgram prog = red_program(p, val, false);
  const std::string key = std::string("red:") + val;
  int use_div = mean ? 1 : 0;
  double factor = mean ? (double)n : 1.0;
  if (pre == 1) return run_reduction_program(p, prog, key, op, RedLayout::Contig, a, 1, oshape, rank, n, slices, 1, use_div, factor, out);
  return run_reduction_program(p, prog, key, op, RedLayout::Strided, a, 1, oshape, rank, n, slices, pre, use_div, factor, out);
}

rm_status reduce_all(rm_provider* p, const rm_handle* a, RedOp op, const char* val, int mean, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "reduce: bad arguments");
  DeviceGuard g(p->ordinal);
  uint64_t one[2] = {1, 1};
  const uint64_t n = handle_elems(a);
  if (n == 0) {
    // simple_provider.rs:6738 (sum of empty = 0), :6893 (mean of empty = 0), max of empty = -inf fold identity
    const double v = op == RedOp::Sum ? 0.0 : op == RedOp::Prod ? 1.0 : op == RedOp::Max ? -INFINITY : INFINITY;
    return rm_fill(p, one, 2, v, out);
  }
  ReductionProgram prog = red_program(p, val, false);
  return run_reduction_program(p, prog, std::string("red:") + val, op, RedLayout::Contig, a, 1, one, 2, n, 1, 1, mean, mean ? (double)n : 1.0, out);
}

}  // namespace

RM_EXPORT rm_status rm_reduce_sum(rm_provider* p, const rm_handle* a, rm_handle* out) { return reduce_all(p, a, RedOp::Sum, "v0", 0, out); }
RM_EXPORT rm_status rm_reduce_prod(rm_provider* p, const rm_handle* a, rm_handle* out) { return reduce_all(p, a, RedOp::Prod, "v0", 0, out); }
RM_EXPORT rm_status rm_reduce_mean(rm_provider* p, const rm_handle* a, rm_handle* out) { return reduce_all(p, a, RedOp::Sum, "v0", 1, out); }
RM_EXPORT rm_status rm_reduce_max(rm_provider* p, const rm_handle* a, rm_handle* out) { return reduce_all(p, a, RedOp::Max, "v0", 0, out); }
RM_EXPORT rm_status rm_reduce_min(rm_provider* p, const rm_handle* a, rm_handle* out) { return reduce_all(p, a, RedOp::Min, "v0", 0, out); }

RM_EXPORT rm_status rm_reduce_sum_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "reduce_sum_dim: bad arguments");
  RM_REQUIRE(dim < std::max<uint32_t>(a->rank, 2), RM_ERROR, "reduce_sum_dim: dim %u out of range", dim);
  DeviceGuard g(p->ordinal);
  return reduce_range(p, a, dim, dim, RedOp::Sum, "v0", 0, out);
}
RM_EXPORT rm_status rm_reduce_mean_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "reduce_mean_dim: bad arguments");
  RM_REQUIRE(dim < std::max<uint32_t>(a->rank, 2), RM_ERROR, "reduce_mean_dim: dim %u out of range", dim);
  DeviceGuard g(p->ordinal);
  return reduce_range(p, a, dim, dim, RedOp::Sum, "v0", 1, out);
}

static rm_status reduce_nd(rm_provider* p, const rm_handle* a, const uint32_t* dims, uint32_t n_dims, const char* val, rm_handle* out) {
  RM_REQUIRE(n_dims > 0, RM_INVALID_ARG, "reduce_mean_nd: no dims");
  std::vector<uint32_t> d(dims, dims + n_dims);
  std::sort(d.begin(), d.end());
  d.erase(std::unique(d.begin(), d.end()), d.end());
  const uint32_t rank = std::max<uint32_t>(a->rank, 2);
  RM_REQUIRE(d.back() < rank, RM_ERROR, "reduce_mean_nd: dim %u out of range", d.back());
  // contiguous groups are reduced in one pass each; the mean divisor is applied group by group
  // (sum(x)/n1/n2 differs from sum(x)/(n1*n2) by <= 1 ulp; the single-group case — the image benchmark's
  // mean(imgs,[2 3]) — is exact).
  rm_handle cur = *a;
  bool owned = false;
  size_t i = 0;
  while (i < d.size()) {
    size_t j = i;
    while (j + 1 < d.size() && d[j + 1] == d[j] + 1) ++j;
    rm_handle next;
    rm_status st = reduce_range(p, &cur, d[i], d[j], RedOp::Sum, i == 0 ? val : "v0", 1, &next);
    std::string msg = st == RM_OK ? "" : last_error();
    if (owned) rm_free(p, &cur);
    if (st != RM_OK) { set_error("%s", msg.c_str()); return st; }
    cur = next;
    owned = true;
    i = j + 1;
  }
  *out = cur;
  return RM_OK;
}

RM_EXPORT rm_status rm_reduce_mean_nd(rm_provider* p, const rm_handle* a, const uint32_t* dims, uint32_t n_dims, rm_handle* out) {
  RM_REQUIRE(p && a && dims && out, RM_INVALID_ARG, "reduce_mean_nd: bad arguments");
  DeviceGuard g(p->ordinal);
  return reduce_nd(p, a, dims, n_dims, "v0", out);
}
RM_EXPORT rm_status rm_reduce_moments_nd(rm_provider* p, const rm_handle* a, const uint32_t* dims, uint32_t n_dims, rm_handle* mean_out, rm_handle* ex2_out) {
  RM_REQUIRE(p && a && dims && mean_out && ex2_out, RM_INVALID_ARG, "reduce_moments_nd: bad arguments");
  DeviceGuard g(p->ordinal);
  RM_TRY(reduce_nd(p, a, dims, n_dims, "v0", mean_out));
  rm_status st = reduce_nd(p, a, dims, n_dims, "(v0 * v0)", ex2_out);
  if (st != RM_OK) { std::string msg = last_error(); rm_free(p, mean_out); set_error("%s", msg.c_str()); }
  return st;
}

// Device-free: NVRTC-compiles the unfused operator surface into the on-disk cubin cache (build() calls this on the
// CPU box so a fresh GPU box does not pay ~0.3 s of NVRTC per first use of each operator).
RM_EXPORT rm_status rm_debug_precompile_ops(rm_precision precision, uint32_t* compiled) {
  rm_provider fake;
  fake.precision = precision;
  uint32_t n = 0;
  std::vector<char> cubin;
  std::string log;
  auto ew = [&](uint32_t n_in, const char* expr, EwVariant v, uint32_t mask) -> rm_status {
    ElementwiseProgram prog = one_node(&fake, n_in, expr);
    RM_TRY(compile_cuda_to_cubin(emit_elementwise_cuda(prog, v, mask), "rm_fused_ew", &cubin, &log));
    ++n;
    return RM_OK;
  };
  for (int op = 0; op < RM_BIN__COUNT; ++op) {
    RM_TRY(ew(2, binary_expr((rm_binary_op)op), EwVariant::Flat, 0));
    RM_TRY(ew(2, binary_expr((rm_binary_op)op), EwVariant::Flat, 2));
    RM_TRY(ew(2, binary_expr((rm_binary_op)op), EwVariant::Broadcast, 0));
  }
  for (int op = 0; op < RM_UN__COUNT; ++op) RM_TRY(ew(1, unary_expr((rm_unary_op)op), EwVariant::Flat, 0));
  for (int op = 0; op < RM_SC__COUNT; ++op) RM_TRY(ew(2, scalar_expr((rm_scalar_op)op), EwVariant::Flat, 2));
  const char* vals[] = {"v0", "(v0 * v0)"};
  for (const char* val : vals)
    for (int op = 0; op < 4; ++op)
      for (int layout = 0; layout < 3; ++layout) {
        if (val != vals[0] && op != 0) continue;
        ReductionProgram prog = red_program(&fake, val, false);
        RM_TRY(compile_cuda_to_cubin(emit_reduction_cuda(prog, (RedOp)op, (RedLayout)layout), "rm_fused_red", &cubin, &log));
        ++n;
      }
  if (compiled) *compiled = n;
  return RM_OK;
}
