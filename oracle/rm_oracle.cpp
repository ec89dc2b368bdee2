// rm_oracle.cpp — CPU restatement of the reference's semantics for the dense-array hot path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path (runmat_b200/, include/) may link, import or
// call this file. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / reported CPU baseline.
//
// Why a restatement: the reference is Rust (toolchain 1.90.0, rust-toolchain.toml:2); cargo/rustc are
// absent in this image and there is no network, so the reference itself cannot be built here
// (SURVEY.md §8c). Every function below cites the reference file:line it follows (paths relative to
// /root/reference/crates). Build: g++ -O2 -ffp-contract=off (Rust never contracts a*b+c to an FMA),
// glibc libm (the libm Rust's f64::{sin,exp,ln,powf,...} resolve to on Linux). Single-threaded, as
// the reference's CPU path is (no rayon/BLAS on it).
//
// Pinning: checked against the literal known-answer vectors in the reference's own unit tests
// (tests/golden/reference_kats.json, each with its file:line) — see tests/test_oracle_kats.py.
// Transcendental *bit* patterns come from libm and are pinned by the reference only to 1e-6..1e-9
// (runmat-accelerate/tests/*.rs); RNG stream values are pinned by algorithm constants, not literals.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

// ---- runmat-runtime/src/builtins/common/broadcast.rs:8-47 (broadcast_shapes) ------------------------
// Pads the SHORTER shape on the FRONT with ones, then per-dim: equal | 1 expands | 0 stays 0 | error.
int broadcast_shapes(const uint64_t* l, int lr, const uint64_t* r, int rr, uint64_t* out) {
  int rank = std::max(lr, rr);
  std::vector<uint64_t> le(rank, 1), re(rank, 1);
  for (int i = 0; i < lr; ++i) le[rank - lr + i] = l[i];
  for (int i = 0; i < rr; ++i) re[rank - rr + i] = r[i];
  for (int d = 0; d < rank; ++d) {
    uint64_t a = le[d], b = re[d];
    if (a == b) out[d] = a;
    else if (a == 1) out[d] = b;
    else if (b == 1) out[d] = a;
    else if (a == 0 || b == 0) out[d] = 0;
    else return -1;
  }
  return rank;
}

// broadcast.rs:50-58 (compute_strides)
void compute_strides(const uint64_t* shape, int rank, uint64_t* strides) {
  uint64_t s = 1;
  for (int i = 0; i < rank; ++i) { strides[i] = s; s *= std::max<uint64_t>(shape[i], 1); }
}

// broadcast.rs:61-92 (broadcast_index). NOTE: uses in_shape.get(dim).unwrap_or(1), i.e. TRAILING
// alignment, whereas broadcast_shapes front-pads. They agree whenever the ranks are equal (always the
// case for MATLAB values, rank >= 2). For unequal ranks we apply the front-padded operand shape so the
// index map is consistent with the output shape; the product does the same (DESIGN.md "broadcast").
uint64_t broadcast_index(uint64_t linear, const uint64_t* out_shape, int out_rank,
                         const uint64_t* in_shape_padded, const uint64_t* strides) {
  uint64_t off = 0;
  for (int d = 0; d < out_rank; ++d) {
    uint64_t oe = out_shape[d];
    uint64_t coord = oe == 0 ? 0 : linear % oe;
    if (oe != 0) linear /= oe;
    uint64_t ie = in_shape_padded[d];
    uint64_t mapped = (ie == 1 || oe == 0) ? 0 : coord;
    off += mapped * strides[d];
  }
  return off;
}

// ---- scalar op semantics -----------------------------------------------------------------------------
// math/rounding/mod.rs:270-300 (mod_real_scalar)
double mod_real_scalar(double a, double b) {
  if (std::isnan(a) || std::isnan(b)) return NAN;
  if (b == 0.0) return NAN;
  if (!std::isfinite(a) && std::isfinite(b)) return NAN;
  double quotient = std::floor(a / b);
  double remainder = a - b * quotient;
  if (remainder == 0.0) remainder = 0.0;
  if (std::isinf(b) && std::isfinite(a)) {
    if (a == 0.0) return 0.0;
    bool sa = std::signbit(a), sb = std::signbit(b);
    return sa == sb ? a : b;
  }
  if (!std::isfinite(remainder) && !std::isfinite(a)) return NAN;
  bool same_sign = remainder == 0.0 || (std::signbit(remainder) == std::signbit(b));
  if (!same_sign) remainder += b;
  if (remainder == 0.0) remainder = 0.0;  // normalises -0.0
  return remainder;
}

// math/rounding/rem.rs:262-281 (rem_real_scalar)
double rem_real_scalar(double a, double b) {
  if (std::isnan(a) || std::isnan(b)) return NAN;
  if (b == 0.0) return NAN;
  if (!std::isfinite(a) && std::isfinite(b)) return NAN;
  if (std::isinf(b) && std::isfinite(a)) return a == 0.0 ? 0.0 : a;
  double q = std::trunc(a / b);
  if (!std::isfinite(q) && std::isfinite(b)) return NAN;
  double r = a - b * q;
  return r == 0.0 ? 0.0 : r;
}

// math/elementwise/sign.rs:236-246
double sign_real_scalar(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : (x == 0 ? 0.0 : x)); }

// math/reduction/max.rs:2323-2343 (choose_real_elementwise, nan_mode = Omit is MATLAB's default for
// max/min(a,b): NaN loses unless both are NaN)
double max_elem(double a, double b) {
  if (std::isnan(a) && std::isnan(b)) return NAN;
  if (std::isnan(a)) return b;
  if (std::isnan(b)) return a;
  return b > a ? b : a;
}
double min_elem(double a, double b) {
  if (std::isnan(a) && std::isnan(b)) return NAN;
  if (std::isnan(a)) return b;
  if (std::isnan(b)) return a;
  return b < a ? b : a;
}

double heaviside(double x) {  // fusion.rs:2944-2952 / runmat-vm/tests/fusion_gpu.rs:2514
  if (std::isnan(x)) return x;
  return x > 0 ? 1.0 : (x == 0 ? 0.5 : 0.0);
}

enum BinOp { B_ADD, B_SUB, B_MUL, B_DIV, B_POW, B_MAX, B_MIN, B_HYPOT, B_ATAN2, B_MOD, B_REM,
             B_GE, B_LE, B_LT, B_GT, B_EQ, B_NE };

double apply_binary(int op, double a, double b) {
  switch (op) {
    case B_ADD: return a + b;                 // plus.rs, simple_provider.rs:481
    case B_SUB: return a - b;                 // simple_provider.rs:482
    case B_MUL: return a * b;                 // times.rs:682-700, simple_provider.rs:483
    case B_DIV: return a / b;                 // simple_provider.rs:484
    case B_POW: return std::pow(a, b);        // power.rs:345-358 (f64::powf), simple_provider.rs:534
    case B_MAX: return max_elem(a, b);
    case B_MIN: return min_elem(a, b);
    case B_HYPOT: return std::hypot(a, b);    // f64::hypot
    case B_ATAN2: return std::atan2(a, b);
    case B_MOD: return mod_real_scalar(a, b);
    case B_REM: return rem_real_scalar(a, b);
    case B_GE: return a >= b ? 1.0 : 0.0;
    case B_LE: return a <= b ? 1.0 : 0.0;
    case B_LT: return a < b ? 1.0 : 0.0;
    case B_GT: return a > b ? 1.0 : 0.0;
    case B_EQ: return a == b ? 1.0 : 0.0;
    case B_NE: return a != b ? 1.0 : 0.0;
  }
  return NAN;
}

enum UnOp { U_SIN, U_COS, U_TAN, U_ASIN, U_ACOS, U_ATAN, U_SINH, U_COSH, U_TANH, U_ASINH, U_ACOSH,
            U_ATANH, U_EXP, U_EXPM1, U_LOG, U_LOG2, U_LOG10, U_LOG1P, U_SQRT, U_ABS, U_SIGN, U_FLOOR,
            U_CEIL, U_ROUND, U_FIX, U_NEG, U_POW2, U_HEAVISIDE, U_SINGLE, U_DOUBLE, U_ISNAN, U_ISINF,
            U_ISFINITE, U_NAN_TO_ZERO, U_NOT_NAN_MASK, U_ERF, U_GAMMA, U_GAMMALN };

double apply_unary(int op, double x) {
  switch (op) {
    case U_SIN: return std::sin(x);     // trigonometry/sin.rs:265-270
    case U_COS: return std::cos(x);
    case U_TAN: return std::tan(x);
    case U_ASIN: return std::asin(x);
    case U_ACOS: return std::acos(x);
    case U_ATAN: return std::atan(x);
    case U_SINH: return std::sinh(x);
    case U_COSH: return std::cosh(x);
    case U_TANH: return std::tanh(x);
    case U_ASINH: return std::asinh(x);
    case U_ACOSH: return std::acosh(x);
    case U_ATANH: return std::atanh(x);
    case U_EXP: return std::exp(x);     // elementwise/exp.rs:170-171
    case U_EXPM1: return std::expm1(x);
    case U_LOG: return std::log(x);
    case U_LOG2: return std::log2(x);
    case U_LOG10: return std::log10(x);
    case U_LOG1P: return std::log1p(x);
    case U_SQRT: return std::sqrt(x);
    case U_ABS: return std::fabs(x);
    case U_SIGN: return sign_real_scalar(x);
    case U_FLOOR: return std::floor(x);
    case U_CEIL: return std::ceil(x);
    case U_ROUND: return std::round(x);  // f64::round = half away from zero
    case U_FIX: return std::trunc(x);
    case U_NEG: return -x;
    case U_POW2: return std::exp2(x);
    case U_HEAVISIDE: return heaviside(x);
    case U_SINGLE: return (double)(float)x;  // times.rs:742-764 rule: value rounded through f32
    case U_DOUBLE: return x;
    case U_ISNAN: return std::isnan(x) ? 1.0 : 0.0;
    case U_ISINF: return std::isinf(x) ? 1.0 : 0.0;
    case U_ISFINITE: return std::isfinite(x) ? 1.0 : 0.0;
    case U_NAN_TO_ZERO: return std::isnan(x) ? 0.0 : x;     // accelerate-api lib.rs:2980
    case U_NOT_NAN_MASK: return std::isnan(x) ? 0.0 : 1.0;  // lib.rs:2985
    case U_ERF: return std::erf(x);       // simple_provider.rs:134-136 (libm crate 0.2.16 erf; glibc erf agrees to <= 1 ulp)
    case U_GAMMA: return std::tgamma(x);
    case U_GAMMALN: return std::lgamma(x);
  }
  return NAN;
}

enum ScOp { S_ADD, S_SUB, S_MUL, S_DIV, S_RSUB, S_RDIV, S_MAX, S_MIN, S_POW };
double apply_scalar(int op, double a, double s) {  // simple_provider.rs:5852-5980
  switch (op) {
    case S_ADD: return a + s;
    case S_SUB: return a - s;
    case S_MUL: return a * s;
    case S_DIV: return a / s;
    case S_RSUB: return s - a;
    case S_RDIV: return s / a;
    case S_MAX: return max_elem(a, s);
    case S_MIN: return min_elem(a, s);
    case S_POW: return std::pow(a, s);
  }
  return NAN;
}

// ---- common/random.rs:9-13, 271-288 -------------------------------------------------------------------
const uint64_t RNG_MULTIPLIER = 6364136223846793005ULL;
const uint64_t RNG_INCREMENT = 1ULL;
const double RNG_SCALE = 1.0 / 9007199254740992.0;  // 1/(1<<53)
const uint64_t DEFAULT_RNG_SEED = 0x9e3779b97f4a7c15ULL;

inline double next_uniform_state(uint64_t* state) {
  *state = *state * RNG_MULTIPLIER + RNG_INCREMENT;
  uint64_t bits = *state >> 11;
  return (double)bits * RNG_SCALE;
}
inline void next_normal_pair(uint64_t* state, double* z0, double* z1) {
  double u1 = next_uniform_state(state);
  if (u1 <= 0.0) u1 = std::numeric_limits<double>::min();  // f64::MIN_POSITIVE
  double u2 = next_uniform_state(state);
  double radius = std::sqrt(-2.0 * std::log(u1));
  double angle = 2.0 * M_PI * u2;
  *z0 = radius * std::cos(angle);
  *z1 = radius * std::sin(angle);
}
void generate_normal(uint64_t* state, uint64_t len, double* out) {  // random.rs:530-543
  uint64_t n = 0;
  while (n < len) {
    double z0, z1;
    next_normal_pair(state, &z0, &z1);
    out[n++] = z0;
    if (n < len) out[n++] = z1;
  }
}

// ---- image/filters/imfilter.rs:750-792 -----------------------------------------------------------------
int64_t clamp_index(int64_t c, int64_t len) { return (len <= 0 || c <= 0) ? 0 : (c >= len ? len - 1 : c); }
int64_t wrap_index(int64_t c, int64_t len) { if (len <= 0) return 0; c %= len; if (c < 0) c += len; return c; }
int64_t reflect_index(int64_t c, int64_t len) {
  if (len <= 0 || len == 1) return 0;
  int64_t period = 2 * len - 2;
  int64_t v = c % period;
  if (v < 0) v += period;
  if (v >= len) v = period - v;
  return v;
}

template <typename T>
void image_normalize_t(const T* data, T* out, uint64_t batch, uint64_t height, uint64_t width, T eps,
                       int has_gain, T gain, int has_bias, T bias, int has_gamma, T gamma,
                       int clamp_zero) {
  // simple_provider.rs:7893-7994. T=double is the host provider verbatim; T=float is the same
  // arithmetic at the f32 precision the wgpu F32 provider (and the benchmark's `single` data) uses.
  uint64_t plane = height * width;
  if (plane == 0) return;
  uint64_t stride_h = batch, stride_w = batch * height;
  for (uint64_t b = 0; b < batch; ++b) {
    T sum = 0;
    for (uint64_t w = 0; w < width; ++w)
      for (uint64_t h = 0; h < height; ++h) sum += data[b + h * stride_h + w * stride_w];
    T mean = sum / (T)plane;
    T sq = 0;
    for (uint64_t w = 0; w < width; ++w)
      for (uint64_t h = 0; h < height; ++h) {
        T d = data[b + h * stride_h + w * stride_w] - mean;
        sq += d * d;
      }
    T variance = sq / (T)plane;
    T sigma = std::sqrt(variance + eps);
    T inv_sigma = sigma > 0 ? (T)1 / sigma : (T)0;
    for (uint64_t w = 0; w < width; ++w)
      for (uint64_t h = 0; h < height; ++h) {
        uint64_t idx = b + h * stride_h + w * stride_w;
        T v = (data[idx] - mean) * inv_sigma;
        if (has_gain) v *= gain;
        if (has_bias) v += bias;
        if (clamp_zero) v = std::fmax(v, (T)0);  // Rust `value.max(0.0)` (simple_provider.rs:7982): NaN.max(0.0) == 0.0, i.e. fmax, not std::max
        if (has_gamma) v = std::pow(v, gamma);
        out[idx] = v;
      }
  }
}

}  // namespace

// ======================================================================================================
// C ABI (ctypes)
// ======================================================================================================

ORC_API int orc_broadcast_shape(const uint64_t* a, int ar, const uint64_t* b, int br, uint64_t* out) {
  return broadcast_shapes(a, ar, b, br, out);
}

// elem_* with MATLAB implicit expansion: simple_provider.rs:459-514 (+ :516-542 for pow).
// `out` must hold prod(broadcast shape) doubles. Returns out rank or -1 on size mismatch.
ORC_API int orc_elem_binary(int op, const double* a, const uint64_t* ashape, int ar, const double* b,
                            const uint64_t* bshape, int br, double* out, uint64_t* out_shape) {
  int rank = broadcast_shapes(ashape, ar, bshape, br, out_shape);
  if (rank < 0) return -1;
  std::vector<uint64_t> ap(rank, 1), bp(rank, 1), as(rank), bs(rank);
  for (int i = 0; i < ar; ++i) ap[rank - ar + i] = ashape[i];
  for (int i = 0; i < br; ++i) bp[rank - br + i] = bshape[i];
  compute_strides(ap.data(), rank, as.data());
  compute_strides(bp.data(), rank, bs.data());
  uint64_t len = 1;
  for (int d = 0; d < rank; ++d) len *= out_shape[d];
  for (uint64_t i = 0; i < len; ++i) {
    uint64_t ia = broadcast_index(i, out_shape, rank, ap.data(), as.data());
    uint64_t ib = broadcast_index(i, out_shape, rank, bp.data(), bs.data());
    out[i] = apply_binary(op, a[ia], b[ib]);
  }
  return rank;
}

ORC_API void orc_unary(int op, const double* a, uint64_t n, double* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = apply_unary(op, a[i]);
}
ORC_API void orc_scalar_op(int op, const double* a, uint64_t n, double s, double* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = apply_scalar(op, a[i], s);
}
ORC_API double orc_mod_scalar(double a, double b) { return mod_real_scalar(a, b); }
ORC_API double orc_rem_scalar(double a, double b) { return rem_real_scalar(a, b); }

// times.rs:742-764: when both operands are single, the f64 result is rounded through f32.
ORC_API void orc_round_through_f32(double* a, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) a[i] = (double)(float)a[i];
}

// The headline fused chain, restated as the sequence of host builtins the VM would run unfused:
// t0 = sin(A) (sin.rs:265-270); t1 = t0 .* B (times.rs:682-700); C = t1 + 1 (plus.rs).
ORC_API void orc_sin_mul_add(const double* a, const double* b, double c, uint64_t n, double* out) {
  for (uint64_t i = 0; i < n; ++i) {
    double t0 = std::sin(a[i]);
    double t1 = t0 * b[i];
    out[i] = t1 + c;
  }
}

// math/reduction/sum.rs:996-1079 (sum_tensor): N-D, sequential column-major accumulation, NaN rules.
// reduce_mask[d] != 0 marks reduced dims. out has prod(shape with reduced dims -> 1) entries.
ORC_API void orc_sum_dims(const double* data, const uint64_t* shape, int rank, const int* reduce_mask,
                          int omit_nan, double* out) {
  std::vector<uint64_t> oshape(rank);
  uint64_t olen = 1, len = 1;
  for (int d = 0; d < rank; ++d) { oshape[d] = reduce_mask[d] ? 1 : shape[d]; olen *= oshape[d]; len *= shape[d]; }
  std::vector<double> sums(olen, 0.0);
  std::vector<char> saw_value(olen, 0), saw_nan(olen, 0);
  std::vector<uint64_t> coords(rank);
  for (uint64_t lin = 0; lin < len; ++lin) {
    uint64_t t = lin;
    for (int d = 0; d < rank; ++d) { coords[d] = shape[d] ? t % shape[d] : 0; if (shape[d]) t /= shape[d]; }
    uint64_t oi = 0, st = 1;
    for (int d = 0; d < rank; ++d) { uint64_t c = reduce_mask[d] ? 0 : coords[d]; oi += c * st; st *= oshape[d]; }
    double v = data[lin];
    if (std::isnan(v)) { if (!omit_nan) saw_nan[oi] = 1; }
    else { sums[oi] += v; saw_value[oi] = 1; }
  }
  for (uint64_t i = 0; i < olen; ++i) {
    if (!omit_nan && saw_nan[i]) out[i] = NAN;
    else out[i] = saw_value[i] ? sums[i] : 0.0;
  }
}

// simple_provider.rs:6728-6746 (reduce_sum = iter().sum()) and :6883-6906 (reduce_mean).
ORC_API double orc_reduce_sum(const double* a, uint64_t n) { double s = 0.0; for (uint64_t i = 0; i < n; ++i) s += a[i]; return s; }
ORC_API double orc_reduce_mean(const double* a, uint64_t n) { return n == 0 ? 0.0 : orc_reduce_sum(a, n) / (double)n; }
ORC_API double orc_reduce_prod(const double* a, uint64_t n) { double s = 1.0; for (uint64_t i = 0; i < n; ++i) s *= a[i]; return s; }
// simple_provider.rs:7366-7385: fold(NEG_INFINITY, f64::max) — f64::max ignores NaN.
ORC_API double orc_reduce_max(const double* a, uint64_t n) { double m = -INFINITY; for (uint64_t i = 0; i < n; ++i) m = std::fmax(m, a[i]); return m; }
ORC_API double orc_reduce_min(const double* a, uint64_t n) { double m = INFINITY; for (uint64_t i = 0; i < n; ++i) m = std::fmin(m, a[i]); return m; }

// simple_provider.rs:7387-7445: 2-D max along dim (0 = over rows), strict '>' so first max wins, NaN
// never replaces; indices 1-based. is_min mirrors reduce_min_dim.
ORC_API void orc_reduce_minmax_dim(const double* a, uint64_t rows, uint64_t cols, int dim, int is_min,
                                   double* vals, double* idx) {
  if (dim == 0) {
    for (uint64_t c = 0; c < cols; ++c) {
      double m = is_min ? INFINITY : -INFINITY; double id = 1.0;
      for (uint64_t r = 0; r < rows; ++r) { double v = a[r + c * rows]; if (is_min ? v < m : v > m) { m = v; id = (double)(r + 1); } }
      vals[c] = m; idx[c] = id;
    }
  } else {
    for (uint64_t r = 0; r < rows; ++r) {
      double m = is_min ? INFINITY : -INFINITY; double id = 1.0;
      for (uint64_t c = 0; c < cols; ++c) { double v = a[r + c * rows]; if (is_min ? v < m : v > m) { m = v; id = (double)(c + 1); } }
      vals[r] = m; idx[r] = id;
    }
  }
}

// builtins/common/linalg.rs:6-32 == simple_provider.rs:7724-7733: C[i+j*m] = sum_k A[i+k*m]*B[k+j*kk],
// k ascending, accumulator starts at 0.0, no FMA.
ORC_API void orc_matmul_naive(const double* a, uint64_t m, uint64_t kk, const double* b, uint64_t n, double* out) {
  for (uint64_t j = 0; j < n; ++j)
    for (uint64_t i = 0; i < m; ++i) {
      double sum = 0.0;
      for (uint64_t k = 0; k < kk; ++k) sum += a[i + k * m] * b[k + j * kk];
      out[i + j * m] = sum;
    }
}
// Loop-interchanged (j,k,i) form: every C[i,j] still receives its products in ascending k starting from
// 0.0, so the result is BIT-IDENTICAL to orc_matmul_naive (tests assert that) while streaming A by
// columns. This is the CPU baseline used at sizes where the strided naive loop is impractical.
ORC_API void orc_matmul(const double* a, uint64_t m, uint64_t kk, const double* b, uint64_t n, double* out) {
  for (uint64_t j = 0; j < n; ++j) {
    double* c = out + j * m;
    for (uint64_t i = 0; i < m; ++i) c[i] = 0.0;
    for (uint64_t k = 0; k < kk; ++k) {
      const double bk = b[k + j * kk];
      const double* ak = a + k * m;
      for (uint64_t i = 0; i < m; ++i) c[i] += ak[i] * bk;
    }
  }
}

// simple_provider.rs:7743-7850: v=acc*alpha+beta; row scale; col scale; clamp_min; clamp_max; powf; diag.
ORC_API void orc_matmul_epilogue(double* c, uint64_t rows, uint64_t cols, double alpha, double beta,
                                 const double* row_scale, int row_div, const double* col_scale, int col_div,
                                 int has_min, double cmin, int has_max, double cmax, int has_pow, double pw,
                                 double* diag) {
  for (uint64_t j = 0; j < cols; ++j)
    for (uint64_t i = 0; i < rows; ++i) {
      uint64_t idx = i + j * rows;
      double v = c[idx] * alpha + beta;
      if (row_scale) v = row_div ? v / row_scale[i] : v * row_scale[i];
      if (col_scale) v = col_div ? v / col_scale[j] : v * col_scale[j];
      if (has_min) v = std::fmax(v, cmin);  // f64::max
      if (has_max) v = std::fmin(v, cmax);
      if (has_pow) v = std::pow(v, pw);
      if (diag && i == j) diag[i] = v;
      c[idx] = v;
    }
}

ORC_API void orc_image_normalize(const double* data, double* out, uint64_t batch, uint64_t height,
                                 uint64_t width, double eps, int has_gain, double gain, int has_bias,
                                 double bias, int has_gamma, double gamma, int clamp_zero) {
  image_normalize_t<double>(data, out, batch, height, width, eps, has_gain, gain, has_bias, bias, has_gamma, gamma, clamp_zero);
}
ORC_API void orc_image_normalize_f32(const float* data, float* out, uint64_t batch, uint64_t height,
                                     uint64_t width, float eps, int has_gain, float gain, int has_bias,
                                     float bias, int has_gamma, float gamma, int clamp_zero) {
  image_normalize_t<float>(data, out, batch, height, width, eps, has_gain, gain, has_bias, bias, has_gamma, gamma, clamp_zero);
}

// image/filters/imfilter.rs:476-745. 2-D/3-D images with a 2-D kernel (rank <= 3). padding: 0 const,
// 1 replicate, 2 symmetric, 3 circular. shape: 0 same, 1 full, 2 valid. mode: 0 corr, 1 conv.
// Returns the output dims in out_shape[3]; `out` may be NULL to query the shape only.
template <typename T>
static int imfilter_t(const T* img, const uint64_t* ishape, int irank, const T* ker,
                      const uint64_t* kshape, int krank, int padding, T cval, int shape, int mode,
                      T* out, uint64_t* out_shape) {
  int rank = std::max(irank, krank);
  if (rank > 3) return -1;
  uint64_t ie[3] = {1, 1, 1}, ke[3] = {1, 1, 1};
  for (int i = 0; i < irank; ++i) ie[i] = ishape[i];
  for (int i = 0; i < krank; ++i) ke[i] = kshape[i];
  int64_t origin[3], base[3]; uint64_t oe[3];
  for (int d = 0; d < 3; ++d) {
    origin[d] = (int64_t)(ke[d] / 2);
    if (shape == 1) { oe[d] = ie[d] + ke[d] - 1; base[d] = origin[d] - ((int64_t)ke[d] - 1); }
    else if (shape == 0) { oe[d] = ie[d]; base[d] = 0; }
    else { oe[d] = ie[d] >= ke[d] ? ie[d] - ke[d] + 1 : 0; base[d] = origin[d]; }
  }
  for (int d = 0; d < 3; ++d) out_shape[d] = oe[d];
  if (!out) return rank;
  uint64_t istr[3] = {1, ie[0], ie[0] * ie[1]};
  uint64_t kstr[3] = {1, ke[0], ke[0] * ke[1]};
  uint64_t ktotal = ke[0] * ke[1] * ke[2];
  uint64_t oi = 0;
  for (uint64_t o2 = 0; o2 < oe[2]; ++o2)
    for (uint64_t o1 = 0; o1 < oe[1]; ++o1)
      for (uint64_t o0 = 0; o0 < oe[0]; ++o0, ++oi) {
        int64_t ob[3] = {(int64_t)o0, (int64_t)o1, (int64_t)o2};
        T sum = (T)0;
        uint64_t kidx[3] = {0, 0, 0};
        for (uint64_t kp = 0; kp < ktotal; ++kp) {  // kernel points in column-major order (:655-690)
          uint64_t lin = kidx[0] * kstr[0] + kidx[1] * kstr[1] + kidx[2] * kstr[2];
          uint64_t flin = (ke[0] - 1 - kidx[0]) * kstr[0] + (ke[1] - 1 - kidx[1]) * kstr[1] + (ke[2] - 1 - kidx[2]) * kstr[2];
          T kv = mode == 0 ? ker[lin] : ker[flin];
          T sample; bool constant = false; uint64_t ilin = 0;
          for (int d = 0; d < 3; ++d) {
            int64_t coord = ob[d] + base[d] + ((int64_t)kidx[d] - origin[d]);
            int64_t len = (int64_t)ie[d];
            if (coord < 0 || coord >= len) {
              if (padding == 0) { constant = true; break; }
              coord = padding == 1 ? clamp_index(coord, len) : (padding == 3 ? wrap_index(coord, len) : reflect_index(coord, len));
            }
            ilin += (uint64_t)coord * istr[d];
          }
          sample = constant ? cval : img[ilin];
          sum += kv * sample;
          for (int d = 0; d < 3; ++d) { if (++kidx[d] < ke[d]) break; kidx[d] = 0; }
        }
        out[oi] = sum;
      }
  return rank;
}

ORC_API int orc_imfilter(const double* img, const uint64_t* ishape, int irank, const double* ker,
                         const uint64_t* kshape, int krank, int padding, double cval, int shape, int mode,
                         double* out, uint64_t* out_shape) {
  return imfilter_t<double>(img, ishape, irank, ker, kshape, krank, padding, cval, shape, mode, out, out_shape);
}
// The same loop in f32 storage arithmetic: what a provider running at ProviderPrecision::F32 computes (the wgpu default,
// backend/wgpu/provider/init.rs:145-172): per-tap `sum += k * sample` rounded to f32 after the multiply and after the add,
// host tap order (imfilter.rs:655-690). Checker for the f32 image config (BASELINE configs[3]).
ORC_API int orc_imfilter_f32(const float* img, const uint64_t* ishape, int irank, const float* ker,
                             const uint64_t* kshape, int krank, int padding, float cval, int shape, int mode,
                             float* out, uint64_t* out_shape) {
  return imfilter_t<float>(img, ishape, irank, ker, kshape, krank, padding, cval, shape, mode, out, out_shape);
}

// conv2d: simple_provider.rs:1845-1876 (conv2d_full_real: scatter over the signal in column-major order, zero
// signal entries skipped, kernel rotated 180 deg) + :1908-1956 (apply_conv2_mode_real_2d). mode: 0 full, 1 same, 2 valid.
ORC_API void orc_conv2d(const double* sig, uint64_t sr, uint64_t sc, const double* ker, uint64_t kr, uint64_t kc, int mode, double* out, uint64_t* out_dims) {
  const uint64_t fr = sr + kr - 1, fc = sc + kc - 1;
  uint64_t r0 = 0, c0 = 0, orows = fr, ocols = fc;
  if (mode == 1) { r0 = (kr - 1) / 2; c0 = (kc - 1) / 2; orows = sr; ocols = sc; }
  else if (mode == 2) { if (sr < kr || sc < kc) { orows = ocols = 0; } else { r0 = kr - 1; c0 = kc - 1; orows = sr - kr + 1; ocols = sc - kc + 1; } }
  out_dims[0] = orows; out_dims[1] = ocols;
  if (!out) return;
  std::vector<double> full(fr * fc, 0.0);
  for (uint64_t c = 0; c < sc; ++c)
    for (uint64_t r = 0; r < sr; ++r) {
      const double aval = sig[c * sr + r];
      if (aval == 0.0) continue;
      for (uint64_t j = 0; j < kc; ++j) {
        const uint64_t oc = c + j, kcol = kc - 1 - j;
        for (uint64_t i = 0; i < kr; ++i) {
          const uint64_t orr = r + i, krow = kr - 1 - i;
          full[oc * fr + orr] += aval * ker[kcol * kr + krow];
        }
      }
    }
  for (uint64_t c = 0; c < ocols; ++c)
    for (uint64_t r = 0; r < orows; ++r) out[c * orows + r] = full[(c0 + c) * fr + r0 + r];
}

// matmul_power_step: simple_provider.rs:7852-7890 — every column of C divided by sqrt(sum(col.^2) + eps), in place.
ORC_API void orc_power_step_normalize(double* c, uint64_t rows, uint64_t cols, double eps) {
  for (uint64_t j = 0; j < cols; ++j) {
    double acc = 0.0;
    for (uint64_t i = 0; i < rows; ++i) { const double v = c[i + j * rows]; acc += v * v; }
    acc += eps;
    const double nrm = std::sqrt(acc);
    for (uint64_t i = 0; i < rows; ++i) c[i + j * rows] /= nrm;
  }
}
// unweighted covariance, rows = All: runmat-runtime/src/builtins/stats/summary/cov.rs:916-960
// (means per column; cov_ij = sum((x_i - m_i)(x_j - m_j)) / denom; denom <= 0 -> NaN; non-finite column -> NaN mean)
ORC_API void orc_covariance(const double* x, uint64_t rows, uint64_t cols, int biased, double* out) {
  const double denom = biased ? (double)rows : (double)rows - 1.0;
  for (uint64_t i = 0; i < cols * cols; ++i) out[i] = NAN;
  if (!(denom > 0.0)) return;
  std::vector<double> means(cols);
  for (uint64_t c = 0; c < cols; ++c) {
    double sum = 0.0; bool valid = true;
    for (uint64_t r = 0; r < rows; ++r) { const double v = x[r + c * rows]; if (!std::isfinite(v)) { valid = false; break; } sum += v; }
    means[c] = valid ? sum / (double)rows : NAN;
  }
  for (uint64_t i = 0; i < cols; ++i)
    for (uint64_t j = i; j < cols; ++j) {
      double acc = 0.0;
      for (uint64_t r = 0; r < rows; ++r) acc += (x[r + i * rows] - means[i]) * (x[r + j * rows] - means[j]);
      out[i + j * cols] = out[j + i * cols] = acc / denom;
    }
}

// ---- RNG: common/random.rs ---------------------------------------------------------------------------------
ORC_API uint64_t orc_default_seed(void) { return DEFAULT_RNG_SEED; }
ORC_API uint64_t orc_mix_seed(uint64_t seed) {  // random.rs:128-142
  if (seed == 0) return DEFAULT_RNG_SEED;
  uint64_t z = seed + 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  uint64_t mixed = z ^ (z >> 31);
  return mixed == 0 ? DEFAULT_RNG_SEED : mixed;
}
ORC_API uint64_t orc_advance_state(uint64_t state, uint64_t delta) {  // random.rs:238-257
  if (delta == 0) return state;
  uint64_t cur_mult = RNG_MULTIPLIER, cur_plus = RNG_INCREMENT, acc_mult = 1, acc_plus = 0;
  while (delta > 0) {
    if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = cur_plus * (cur_mult + 1);
    cur_mult = cur_mult * cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}
ORC_API uint64_t orc_generate_uniform(uint64_t state, uint64_t n, double* out) {  // simple_provider.rs:3514-3533
  for (uint64_t i = 0; i < n; ++i) out[i] = next_uniform_state(&state);
  return state;
}
ORC_API uint64_t orc_generate_normal(uint64_t state, uint64_t n, double* out) {  // random.rs:530-543
  generate_normal(&state, n, out);
  return state;
}

// stats/random/stochastic_evolution.rs:11-32: `steps` sequential passes; each pass draws len normals
// from the global LCG (pairs: element 2j gets r*cos, 2j+1 gets r*sin) and multiplies by exp(drift+scale*z).
// Returns the RNG state after the run.
ORC_API uint64_t orc_stochastic_evolution(uint64_t rng_state, double* data, uint64_t len, double drift,
                                          double scale, uint32_t steps) {
  if (len == 0 || steps == 0) return rng_state;
  std::vector<double> samples(len);
  for (uint32_t t = 0; t < steps; ++t) {
    generate_normal(&rng_state, len, samples.data());
    for (uint64_t i = 0; i < len; ++i) {
      double term = drift + scale * samples[i];
      data[i] *= std::exp(term);
    }
  }
  return rng_state;
}

// The same host algorithm replayed for a SAMPLE of paths (full-size parity of the 1e8 x 256 run: the whole vector would be
// 2.56e10 sequential draws). In pass t the host's generate_normal consumes 2 draws per Box-Muller pair, pair j = path/2 first
// (random.rs:530-543), so path p's normal of pass t starts at stream offset t*2*ceil(len/2) + 2*(p/2); orc_advance_state
// (random.rs:238-257, pinned against n sequential steps in tests/test_oracle_kats.py) jumps there. Element 2j takes
// r*cos, element 2j+1 takes r*sin; value *= exp(drift + scale*z) with a rounding after every operation, as in :25-27.
ORC_API void orc_stochastic_evolution_sampled(uint64_t rng_state, double s0, uint64_t len, double drift, double scale, uint32_t steps,
                                              const uint64_t* paths, uint64_t n_paths, double* out) {
  const uint64_t draws_per_pass = 2 * ((len + 1) / 2);
  for (uint64_t k = 0; k < n_paths; ++k) {
    const uint64_t p = paths[k];
    double v = s0;
    for (uint32_t t = 0; t < steps; ++t) {
      uint64_t st = orc_advance_state(rng_state, (uint64_t)t * draws_per_pass + 2 * (p / 2));
      double z0, z1;
      next_normal_pair(&st, &z0, &z1);
      double term = drift + scale * ((p & 1) ? z1 : z0);
      v *= std::exp(term);
    }
    out[k] = v;
  }
}

// builtins/math/linalg/solve/linsolve.rs:769-833 (forward_substitution_real / backward_substitution_real): per right-hand
// side, accum = sum_j T[i,j]*x[j] in ascending j, x[i] = (b[i] - accum) / T[i,i]; rcond = min|diag| / max|diag|
// (diagonal_rcond, common/linalg.rs:232-238). Returns -1 on a zero diagonal entry ("singular to working precision").
ORC_API int orc_linsolve_triangular(const double* t, uint64_t n, const double* rhs, uint64_t nrhs, int lower, double* out, double* rcond) {
  double min_diag = INFINITY, max_diag = 0.0;
  for (uint64_t i = 0; i < n * nrhs; ++i) out[i] = rhs[i];
  for (uint64_t col = 0; col < nrhs; ++col) {
    for (uint64_t step = 0; step < n; ++step) {
      const uint64_t i = lower ? step : n - 1 - step;
      const double diag = t[i + i * n], diag_abs = std::fabs(diag);
      min_diag = std::fmin(min_diag, diag_abs);
      max_diag = std::fmax(max_diag, diag_abs);
      if (diag_abs == 0.0) return -1;
      double accum = 0.0;
      if (lower) { for (uint64_t j = 0; j < i; ++j) accum += t[i + j * n] * out[j + col * n]; }
      else { for (uint64_t j = i + 1; j < n; ++j) accum += t[i + j * n] * out[j + col * n]; }
      out[i + col * n] = (out[i + col * n] - accum) / diag;
    }
  }
  *rcond = max_diag == 0.0 ? 0.0 : min_diag / max_diag;
  return 0;
}

// simple_provider.rs:3488-3512 (linspace; last element forced to `stop`)
ORC_API void orc_linspace(double start, double stop, uint64_t count, double* out) {
  if (count == 0) return;
  if (count == 1) { out[0] = stop; return; }
  double step = (stop - start) / (double)(count - 1);
  for (uint64_t i = 0; i < count; ++i) out[i] = start + (double)i * step;
  out[count - 1] = stop;
}

// simple_provider.rs:5983- (transpose, 2-D column-major)
ORC_API void orc_transpose(const double* a, uint64_t rows, uint64_t cols, double* out) {
  for (uint64_t c = 0; c < cols; ++c)
    for (uint64_t r = 0; r < rows; ++r) out[c + r * cols] = a[r + c * rows];
}
// simple_provider.rs:2609-2653 / :2655-2713 (gather_linear / scatter_linear); -1 on out-of-bounds.
ORC_API int orc_gather_linear(const double* src, uint64_t n, const uint32_t* idx, uint64_t ni, double* out) {
  for (uint64_t i = 0; i < ni; ++i) { if (idx[i] >= n) return -1; out[i] = src[idx[i]]; }
  return 0;
}
ORC_API int orc_scatter_linear(double* dst, uint64_t n, const uint32_t* idx, uint64_t ni, const double* vals) {
  for (uint64_t i = 0; i < ni; ++i) if (idx[i] >= n) return -1;
  for (uint64_t i = 0; i < ni; ++i) dst[idx[i]] = vals[i];
  return 0;
}

// ---- benchmark-level restatements (the .m scripts' arithmetic, evaluated with host builtin semantics) ---
// benchmarks/monte-carlo-analysis/runmat_lcg.m:33-51. S is `single`: every op whose operands are single
// is rounded through f32 (times.rs:742-764); the LCG index arithmetic is explicit double.
// Evolves paths [path0, path0+count) of M and returns sum(max(S-K,0)) over them (double accumulation of
// the single payoffs); price = total/M*exp(-mu*T*dt) is formed by the caller.
ORC_API double orc_mc_lcg_payoff_sum(uint64_t M, uint32_t T, uint64_t path0, uint64_t count, double seed,
                                     float S0, float mu, float sigma, float dt, float K, float* s_out) {
  float sqrt_dt = std::sqrt(dt);
  float drift = (mu - 0.5f * (sigma * sigma)) * dt;
  float scale = sigma * sqrt_dt;
  double twoM = (double)M * 2.0;
  double total = 0.0;
  for (uint64_t p = path0; p < path0 + count; ++p) {
    float S = 1.0f * S0;
    for (uint32_t t = 0; t < T; ++t) {
      double salt = (double)t * twoM;
      double idx1 = (double)p + salt + seed;
      double idx2 = (double)p + salt + (double)M + seed;
      double state1 = mod_real_scalar(1664525.0 * idx1 + 1013904223.0, 4294967296.0);
      double state2 = mod_real_scalar(1664525.0 * idx2 + 1013904223.0, 4294967296.0);
      double u1 = std::fmax(state1 / 4294967296.0, 1.0 / 4294967296.0);
      double u2 = state2 / 4294967296.0;
      double r = std::sqrt(-2.0 * std::log(u1));
      double theta = 2.0 * M_PI * u2;
      float z = (float)(r * std::cos(theta));
      float e = (float)std::exp((double)(drift + (float)(scale * z)));  // single exp: f64 libm rounded to f32
      S = S * e;
    }
    if (s_out) s_out[p - path0] = S;
    float payoff = std::fmax(S - K, 0.0f);
    total += (double)payoff;
  }
  return total;
}

// benchmarks/4k-image-processing/runmat_lcg.m:59-79: imgs(b,h,w) = single(mod(1664525*idx+1013904223,2^32))/single(2^32),
// idx = (b-1)*H*W + seed + h*W + w (0-based h,w); layout [B,H,W] column-major (batch stride 1).
ORC_API void orc_image_lcg_fill(float* imgs, uint64_t B, uint64_t H, uint64_t W, double seed, uint64_t b0, uint64_t bcount) {
  // fills a [bcount,H,W] tensor holding global images b0..b0+bcount-1
  for (uint64_t w = 0; w < W; ++w)
    for (uint64_t h = 0; h < H; ++h)
      for (uint64_t b = 0; b < bcount; ++b) {
        double idx = (double)((b0 + b) * H * W) + seed + (double)h * (double)W + (double)w;
        double state = mod_real_scalar(1664525.0 * idx + 1013904223.0, 4294967296.0);
        imgs[b + h * bcount + w * bcount * H] = (float)state / 4294967296.0f;
      }
  (void)B;
}
// runmat_lcg.m:86-100: mse = mean((out - imgs).^2,'all') in single, accumulated here in double for a
// stable checker value (the device result is compared with a relative tolerance).
ORC_API double orc_sq_err_sum_f32(const float* a, const float* b, uint64_t n) {
  double s = 0.0;
  for (uint64_t i = 0; i < n; ++i) { float e = a[i] - b[i]; s += (double)(e * e); }
  return s;
}
