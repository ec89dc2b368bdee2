// fusion_lower.cpp — lowers the reference planner's fused programs to CUDA C for sm_100a.
//
// The reference hands a provider WGSL *text* (AccelProvider::fused_elementwise / fused_reduction,
// crates/runmat-accelerate-api/src/lib.rs:2946-3007). The text is produced by a closed generator
// (crates/runmat-accelerate/src/fusion.rs:1525-1763 for elementwise, :1765-2077 for reductions): a fixed
// prologue, then one `let tmpN: T = <expr>;` per fused op and `output[k].data[g] = <expr>;` stores, with
// <expr> drawn from the primitive_expr / builtin_expr tables (fusion.rs:2874-3026). The reference's own
// test provider re-parses exactly this subset (crates/runmat-vm/tests/fusion_gpu.rs:886-1013). We do the
// same and emit a CUDA kernel built around hand-written load/store/reduce scaffolding:
//   * 256-bit (LDG.E.256 / STG.E.256) coalesced streaming loads and stores, several in flight per thread,
//   * scalars (1-element inputs, how the executor passes constants: fusion_exec.rs:305-326) hoisted,
//   * a general broadcast variant driven by coalesced dims/strides,
//   * reductions: per-thread accumulate -> warp shuffle -> shared -> deterministic last-block finish.
// Arithmetic follows the CPU builtins' rounding: compiled with -fmad=false (Rust never fuses a*b+c).
#include <cctype>
#include <cstdlib>
#include <sstream>

#include "common.h"

namespace rm {

namespace {

struct Tok {
  enum Kind { Ident, Number, Punct, End } kind;
  std::string text;
};

bool tokenize(const std::string& s, std::vector<Tok>* out, std::string* err) {
  size_t i = 0, n = s.size();
  while (i < n) {
    char c = s[i];
    if (isspace((unsigned char)c)) { ++i; continue; }
    if (isalpha((unsigned char)c) || c == '_') {
      size_t j = i + 1;
      while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j;
      out->push_back({Tok::Ident, s.substr(i, j - i)});
      i = j;
      continue;
    }
    if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)s[i + 1]))) {
      size_t j = i;
      while (j < n && (isdigit((unsigned char)s[j]) || s[j] == '.')) ++j;
      if (j < n && (s[j] == 'e' || s[j] == 'E')) {
        size_t k = j + 1;
        if (k < n && (s[k] == '+' || s[k] == '-')) ++k;
        if (k < n && isdigit((unsigned char)s[k])) {
          j = k;
          while (j < n && isdigit((unsigned char)s[j])) ++j;
        }
      }
      std::string num = s.substr(i, j - i);
      // WGSL suffixes: u (u32), i (i32), f (f32), h (f16)
      if (j < n && (s[j] == 'u' || s[j] == 'i' || s[j] == 'f' || s[j] == 'h')) {
        if (s[j] == 'u' || s[j] == 'i') { num += (s[j] == 'u' ? "u" : ""); }
        ++j;
      }
      out->push_back({Tok::Number, num});
      i = j;
      continue;
    }
    // two-char operators
    if (i + 1 < n) {
      std::string two = s.substr(i, 2);
      if (two == "&&" || two == "||" || two == "==" || two == "!=" || two == "<=" || two == ">=") {
        out->push_back({Tok::Punct, two});
        i += 2;
        continue;
      }
    }
    if (strchr("+-*/()<>!,.[]%", c)) {
      out->push_back({Tok::Punct, std::string(1, c)});
      ++i;
      continue;
    }
    *err = std::string("unexpected character '") + c + "' in fused expression";
    return false;
  }
  out->push_back({Tok::End, ""});
  return true;
}

// WGSL builtin -> CUDA device function (prelude below defines the rm_* ones).
const char* map_function(const std::string& id) {
  static const std::unordered_map<std::string, const char*> table = {
      {"sin", "rm_sin"},  {"cos", "rm_cos"},     {"tan", "tan"},     {"asin", "asin"},   {"acos", "acos"},
      {"atan", "atan"},   {"atan2", "atan2"}, {"sinh", "sinh"},   {"cosh", "cosh"},   {"tanh", "tanh"},
      {"asinh", "asinh"}, {"acosh", "acosh"}, {"atanh", "atanh"}, {"exp", "exp"},     {"exp2", "exp2"},
      {"log", "log"},     {"log2", "log2"},   {"sqrt", "sqrt"},   {"abs", "fabs"},    {"floor", "floor"},
      {"ceil", "ceil"},   {"round", "round"}, {"trunc", "trunc"}, {"sign", "rm_sign"}, {"pow", "pow"},
      {"max", "fmax"},    {"min", "fmin"},    {"hypot", "hypot"}, {"select", "rm_select"},
      {"isNan", "rm_isnan"}, {"isInf", "rm_isinf"}, {"isFinite", "rm_isfinite"}, {"isNanF", "rm_isnan"},
      {"f64", "rm_f64"},  {"f32", "rm_f32"},
  };
  auto it = table.find(id);
  return it == table.end() ? nullptr : it->second;
}

bool is_value_ident(const std::string& id) {
  // tmpN (elementwise temporaries), v / vN (reduction operands, and our rewritten inputs)
  if (id.rfind("tmp", 0) == 0 && id.size() > 3) {
    for (size_t i = 3; i < id.size(); ++i) if (!isdigit((unsigned char)id[i])) return false;
    return true;
  }
  if (id == "v") return true;
  if (id[0] == 'v' && id.size() > 1) {
    for (size_t i = 1; i < id.size(); ++i) if (!isdigit((unsigned char)id[i])) return false;
    return true;
  }
  return false;
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

bool parse_scalar_ty(const std::string& sh, std::string* ty, std::string* err) {
  size_t p = sh.find("struct Tensor");
  if (p != std::string::npos) p = sh.find("array<", p);
  if (p == std::string::npos) { *err = "fused shader: missing `struct Tensor { data: array<T> }`"; return false; }
  size_t q = sh.find('>', p);
  *ty = trim(sh.substr(p + 6, q - (p + 6)));
  if (*ty != "f64" && *ty != "f32") { *err = "fused shader: unsupported scalar type " + *ty; return false; }
  return true;
}

uint32_t count_inputs(const std::string& sh) {
  uint32_t n = 0;
  for (;;) {
    std::string needle = "var<storage, read> input" + std::to_string(n) + ":";
    if (sh.find(needle) == std::string::npos) break;
    ++n;
  }
  return n;
}

}  // namespace

bool translate_expr(const std::string& wgsl, const std::string& scalar_ty, std::string* cuda, std::string* err) {
  std::vector<Tok> toks;
  if (!tokenize(wgsl, &toks, err)) return false;
  std::string out;
  const bool f32 = scalar_ty == "f32";
  for (size_t i = 0; i + 1 < toks.size(); ++i) {
    const Tok& t = toks[i];
    if (t.kind == Tok::Ident) {
      // inputK.data[iK]  ->  vK   (generator: fusion.rs:1650 / :1719)
      if (t.text.rfind("input", 0) == 0 && i + 6 < toks.size() && toks[i + 1].text == "." &&
          toks[i + 2].text == "data" && toks[i + 3].text == "[" && toks[i + 5].text == "]") {
        std::string k = t.text.substr(5);
        bool digits = !k.empty();
        for (char ch : k) digits = digits && isdigit((unsigned char)ch);
        if (!digits || toks[i + 4].text != "i" + k) { *err = "fused expression: malformed input reference near " + t.text; return false; }
        out += "v" + k;
        i += 5;
        continue;
      }
      const bool is_call = toks[i + 1].kind == Tok::Punct && toks[i + 1].text == "(";
      if (is_call) {
        const char* fn = map_function(t.text);
        if (!fn) { *err = "fused expression: unsupported function `" + t.text + "`"; return false; }
        out += fn;
        continue;
      }
      if (is_value_ident(t.text)) { out += (t.text == "v" ? std::string("v0") : t.text); continue; }
      if (t.text == "true" || t.text == "false") { out += t.text; continue; }
      *err = "fused expression: unknown identifier `" + t.text + "`";
      return false;
    }
    if (t.kind == Tok::Number) {
      bool is_float = t.text.find('.') != std::string::npos || t.text.find('e') != std::string::npos ||
                      t.text.find('E') != std::string::npos;
      bool is_uint = !t.text.empty() && t.text.back() == 'u';
      if (is_uint) { *err = "fused expression: unexpected integer literal " + t.text; return false; }
      // abstract-float literals take the scalar type (WGSL); keep integers exact by spelling them as floats
      out += t.text;
      if (!is_float) out += ".0";
      if (f32) out += "f";
      continue;
    }
    if (t.text == "." || t.text == "[" || t.text == "]" || t.text == "%") {
      *err = "fused expression: unsupported token `" + t.text + "`";
      return false;
    }
    out += t.text;
    if (t.text == ",") out += " ";
  }
  *cuda = out;
  return true;
}

bool parse_elementwise_wgsl(const char* shader, ElementwiseProgram* prog, std::string* err) {
  std::string sh(shader ? shader : "");
  if (!parse_scalar_ty(sh, &prog->scalar_ty, err)) return false;
  prog->n_inputs = count_inputs(sh);
  if (prog->n_inputs == 0) { *err = "fused_elementwise: no inputs"; return false; }  // elementwise.rs:1575
  size_t body = sh.find("fn main(");
  if (body == std::string::npos) { *err = "fused_elementwise: missing entry point"; return false; }
  std::istringstream lines(sh.substr(body));
  std::string line;
  std::vector<std::pair<int, std::string>> outs;
  while (std::getline(lines, line)) {
    std::string t = trim(line);
    if (t.rfind("let tmp", 0) == 0) {
      size_t colon = t.find(':'), eq = t.find('=');
      if (colon == std::string::npos || eq == std::string::npos || t.back() != ';') { *err = "fused_elementwise: malformed statement: " + t; return false; }
      std::string name = trim(t.substr(4, colon - 4));
      std::string expr = trim(t.substr(eq + 1, t.size() - eq - 2));
      std::string cu;
      if (!translate_expr(expr, prog->scalar_ty, &cu, err)) return false;
      prog->stmts.push_back("const T " + name + " = " + cu + ";");
    } else if (t.rfind("output", 0) == 0 && t.find(".data[g]") != std::string::npos) {
      size_t dot = t.find(".data[g]"), eq = t.find('=', dot);
      std::string idx = t.substr(6, dot - 6);
      int k = idx.empty() ? 0 : atoi(idx.c_str());
      std::string expr = trim(t.substr(eq + 1, t.size() - eq - 2));
      std::string cu;
      if (!translate_expr(expr, prog->scalar_ty, &cu, err)) return false;
      outs.push_back({k, cu});
    }
  }
  if (outs.empty()) { *err = "fused_elementwise: shader has no output store"; return false; }
  prog->n_outputs = (uint32_t)outs.size();
  prog->outputs.assign(outs.size(), "");
  for (auto& o : outs) {
    if (o.first < 0 || (size_t)o.first >= outs.size()) { *err = "fused_elementwise: bad output index"; return false; }
    prog->outputs[o.first] = o.second;
  }
  return true;
}

bool parse_reduction_wgsl(const char* shader, ReductionProgram* prog, std::string* err) {
  std::string sh(shader ? shader : "");
  if (!parse_scalar_ty(sh, &prog->scalar_ty, err)) return false;
  prog->n_inputs = count_inputs(sh);
  if (prog->n_inputs == 0) { *err = "fused_reduction: no inputs"; return false; }
  prog->omit_nan = sh.find("const OMITNAN: bool = true") != std::string::npos;
  if (sh.find("let col = wid.x;") != std::string::npos) prog->axis = 0;       // fusion.rs:1976
  else if (sh.find("let row = wid.x;") != std::string::npos) prog->axis = 1;  // fusion.rs:2034
  else { *err = "fused_reduction: cannot determine reduction axis"; return false; }
  size_t p = sh.find("let val:");
  if (p == std::string::npos) { *err = "fused_reduction: missing `let val` operand"; return false; }
  size_t eq = sh.find('=', p), semi = sh.find(';', eq);
  std::string expr = trim(sh.substr(eq + 1, semi - eq - 1));
  return translate_expr(expr, prog->scalar_ty, &prog->val_expr, err);
}

// -----------------------------------------------------------------------------------------------------------
// CUDA emission
// -----------------------------------------------------------------------------------------------------------
namespace {

// Hand-written device scaffolding shared by every generated kernel.
const char* kPrelude = R"CUDA(
typedef unsigned long long u64;
typedef unsigned int u32;
#if RM_F32
typedef float T;
#define VEC 8
#else
typedef double T;
#define VEC 4
#endif
struct __align__(32) vec_t { T x[VEC]; };
// Programmatic dependent launch (fused.cu launches with programmatic stream serialisation): let the NEXT kernel in the stream
// become resident as soon as every CTA of this one has started, then wait here until the PREVIOUS kernel has completed and its
// writes are visible. No global memory is touched before the wait, so stream order is preserved exactly; what overlaps is the
// previous kernel's drain with this kernel's launch latency and ramp. Both instructions are no-ops in a plain launch.
#define RM_PDL_PROLOGUE() do { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); } while (0)
#define RM_NAN __longlong_as_double(0x7ff8000000000000LL)

// 256-bit streaming accesses: LDG.E.256 / STG.E.256 on sm_100a, L1 no-allocate (each byte is touched once).
__device__ __forceinline__ vec_t ldv(const T* p) {
  vec_t r;
#if RM_F32
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.x[0]), "=f"(r.x[1]), "=f"(r.x[2]), "=f"(r.x[3]), "=f"(r.x[4]), "=f"(r.x[5]), "=f"(r.x[6]), "=f"(r.x[7])
               : "l"(p));
#else
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x[0]), "=d"(r.x[1]), "=d"(r.x[2]), "=d"(r.x[3]) : "l"(p));
#endif
  return r;
}
__device__ __forceinline__ void stv(T* p, const vec_t& r) {
#if RM_F32
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(r.x[0]), "f"(r.x[1]), "f"(r.x[2]), "f"(r.x[3]), "f"(r.x[4]), "f"(r.x[5]), "f"(r.x[6]), "f"(r.x[7]) : "memory");
#else
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(r.x[0]), "d"(r.x[1]), "d"(r.x[2]), "d"(r.x[3]) : "memory");
#endif
}

// WGSL helpers (fusion.rs:1559-1590) with the CPU builtins' semantics.
__device__ __forceinline__ T rm_select(T f, T t, bool c) { return c ? t : f; }   // select(f, t, cond)
__device__ __forceinline__ bool rm_isnan(T x) { return x != x; }
__device__ __forceinline__ bool rm_isinf(T x) { return isinf(x); }
__device__ __forceinline__ bool rm_isfinite(T x) { return isfinite(x); }
__device__ __forceinline__ T rm_sign(T x) { return x > (T)0 ? (T)1 : (x < (T)0 ? (T)-1 : x); }  // sign.rs:236-246 (+-0 and NaN pass through)
// math/rounding/mod.rs:270-300 (mod_real_scalar) and rem.rs:262-281 (rem_real_scalar)
__device__ __forceinline__ T rm_mod(T a, T b) {
  if (a != a || b != b || b == (T)0) return (T)RM_NAN;
  if (!isfinite(a) && isfinite(b)) return (T)RM_NAN;
  const T q = floor(a / b);
  T r = a - b * q;
  if (isinf(b) && isfinite(a)) { if (a == (T)0) return (T)0; return (signbit(a) == signbit(b)) ? a : b; }
  if (!isfinite(r) && !isfinite(a)) return (T)RM_NAN;
  if (!(r == (T)0 || signbit(r) == signbit(b))) r += b;
  if (r == (T)0) r = (T)0;
  return r;
}
__device__ __forceinline__ T rm_rem(T a, T b) {
  if (a != a || b != b || b == (T)0) return (T)RM_NAN;
  if (!isfinite(a) && isfinite(b)) return (T)RM_NAN;
  if (isinf(b) && isfinite(a)) return a == (T)0 ? (T)0 : a;
  const T q = trunc(a / b);
  if (!isfinite(q) && isfinite(b)) return (T)RM_NAN;
  const T r = a - b * q;
  return r == (T)0 ? (T)0 : r;
}
__device__ __forceinline__ T rm_heaviside(T x) { return x != x ? x : (x > (T)0 ? (T)1 : (x == (T)0 ? (T)0.5 : (T)0)); }
__device__ __forceinline__ double rm_f64(double x) { return x; }
__device__ __forceinline__ float rm_f32(double x) { return (float)x; }
#if RM_F32 || RM_LIBM_TRIG
#define rm_sin(x) sin(x)
#define rm_cos(x) cos(x)
#else
// ---- rm_trig begin (tests/test_lowering.py compiles this block for the host and compares it with glibc) ----------------------
// f64 sin / cos for the generated kernels. r47 ncu + SASS of the headline reduction sum(sin(A).*B+1): 62 instructions per element
// of which 17 go to the FP64 pipe; the rest is the CUDA library's quadrant logic (two coefficient sets fetched from a table with
// three LDG.128, FSEL pairs, F2I/I2F, UMOV pairs that rebuild immediates): issue slots 59 % busy, and the kernel runs at 50 us
// where the same chain without the sin runs at 42 us (profiles/r49_red_sweep.txt). Here the argument is reduced modulo pi
// instead of pi/2 (three-part pi, the first two parts 33 bits wide so q * part is exact for |q| < 2^20), ONE odd minimax
// polynomial covers |r| <= pi/2 (degree 17, 3.9e-17 relative; Chebyshev-node fit in 60-digit arithmetic), the quadrant is a sign
// flip on the integer side, and every constant is a __constant__ bank operand of its DFMA. cos uses half-integer multiples of pi,
// so it keeps full relative accuracy next to its zeros as well. ~25 instructions, 15 FP64. Measured against glibc on the host:
// <= 1.5 ulp over the fast range. |x| < 2^-27 returns the fdlibm answer (x, 1); |x| >= 105615, Inf and NaN take the library
// routine out of line. RUNMAT_B200_LIBM_TRIG=1 restores the library functions.
__constant__ double rm_kTrig[14] = {
  0x1.45f306dc9c883p-2, 6755399441055744.0,                                      // 1/pi, 1.5 * 2^52
  -0x1.921fb54400000p+1, -0x1.0b4611a600000p-33, -0x1.3198a2e037073p-68,         // -pi in three parts
  0x1.899d1babb713bp-49, -0x1.ae50fd31b6504p-41, 0x1.612400a095576p-33, -0x1.ae64559f961c0p-26,
  0x1.71de3a5452e63p-19, -0x1.a01a01a018a59p-13, 0x1.1111111111107p-7, -0x1.5555555555555p-3,
  0.5};
__device__ __noinline__ double rm_sin_slow(double x) { return sin(x); }
__device__ __noinline__ double rm_cos_slow(double x) { return cos(x); }
// sin(x - h*pi) with the sign of bit 0 of `flip` applied; |x - h*pi| <= pi/2 (+ rounding slop)
__device__ __forceinline__ double rm_sin_reduced(double x, double h, int flip) {
  double r = fma(h, rm_kTrig[2], x);
  r = fma(h, rm_kTrig[3], r);
  r = fma(h, rm_kTrig[4], r);
  const double z = r * r;
  double p = rm_kTrig[5];
  #pragma unroll
  for (int i = 6; i < 13; ++i) p = fma(p, z, rm_kTrig[i]);
  const double s = fma(r * z, p, r);
  return __hiloint2double(__double2hiint(s) ^ (flip << 31), __double2loint(s));
}
__device__ __forceinline__ double rm_sin(double x) {
  const u32 hx = (u32)__double2hiint(x) & 0x7fffffffu;
  if (hx - 0x3e400000u >= 0x40f9c8f0u - 0x3e400000u) return hx < 0x3e400000u ? x : rm_sin_slow(x);
  const double t = fma(x, rm_kTrig[0], rm_kTrig[1]);   // round(x / pi) in the low mantissa bits
  return rm_sin_reduced(x, t - rm_kTrig[1], __double2loint(t));
}
__device__ __forceinline__ double rm_cos(double x) {
  const u32 hx = (u32)__double2hiint(x) & 0x7fffffffu;
  if (hx - 0x3e400000u >= 0x40f9c8f0u - 0x3e400000u) return hx < 0x3e400000u ? 1.0 : rm_cos_slow(x);
  const double t = fma(x, rm_kTrig[0], -rm_kTrig[13]) + rm_kTrig[1];   // n = round(x / pi - 1/2)
  // cos(x) = cos(r + (n + 1/2) pi) = (-1)^(n+1) sin(r)
  return rm_sin_reduced(x, (t - rm_kTrig[1]) + rm_kTrig[13], __double2loint(t) + 1);
}
// ---- rm_trig end
#endif
)CUDA";

// tuning knobs (defaults chosen from the r01 sweep on B200; overridable for experiments)
int env_int(const char* name, int dflt, int lo, int hi) {
  if (const char* e = getenv(name)) { int v = atoi(e); if (v >= lo && v <= hi) return v; }
  return dflt;
}
int libm_trig() { return env_int("RUNMAT_B200_LIBM_TRIG", 0, 0, 1); }
int red_unroll() { return env_int("RUNMAT_B200_RED_UNROLL", 2, 1, 8); }
// 4 resident CTAs/SM (<= 64 registers) + two independent accumulators: 49.2 us vs 53.9 us for the r01 structure on the headline
// reduction (profiles/r04_harness_results.txt: occupancy is the lever, a third/fourth accumulator costs registers and loses)
// Contig reductions sweep their input from the END to the front. The elementwise kernels sweep front to end, so in the common
// sequence "C = f(A, B); s = sum(g(A, B))" (the benchmark's step) the reduction starts on the ~40 % of A and B that the previous
// kernel left in the 126 MB L2, and it finishes at the front, where the next forward sweep starts. Any fixed order is
// deterministic; RUNMAT_B200_RED_FORWARD=1 restores the forward sweep.
int red_reverse() { return env_int("RUNMAT_B200_RED_FORWARD", 0, 0, 1) ? 0 : 1; }
int red_blocked() { return env_int("RUNMAT_B200_RED_BLOCKED", 0, 0, 1); }
int red_minblocks() { return env_int("RUNMAT_B200_RED_MINB", 4, 0, 8); }

std::string input_params(uint32_t n_inputs) {
  std::string s;
  for (uint32_t k = 0; k < n_inputs; ++k) s += "const T* __restrict__ in" + std::to_string(k) + ", ";
  return s;
}

}  // namespace

std::string emit_elementwise_cuda(const ElementwiseProgram& prog, EwVariant variant, uint32_t scalar_mask) {
  std::ostringstream o;
  o << "#define RM_F32 " << (prog.scalar_ty == "f32" ? 1 : 0) << "\n#define RM_LIBM_TRIG " << libm_trig() << "\n" << kPrelude;
  const uint32_t ni = prog.n_inputs, no = prog.n_outputs;
  std::string body;
  for (auto& s : prog.stmts) body += "      " + s + "\n";

  if (variant == EwVariant::Flat) {
    // One thread: UNROLL 256-bit vectors per streamed input, all loads issued before any math.
    o << "#define UNROLL 2\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256) rm_fused_ew(" << input_params(ni);
    for (uint32_t k = 0; k < no; ++k) o << "T* __restrict__ out" << k << ", ";
    o << "u64 n) {\n  RM_PDL_PROLOGUE();\n";
    o << "  const u64 nvec = n / VEC;\n";
    for (uint32_t k = 0; k < ni; ++k)
      if (scalar_mask & (1u << k)) o << "  const T s" << k << " = in" << k << "[0];\n";
    o << "  const u64 base = (u64)blockIdx.x * (UNROLL * 256) + threadIdx.x;\n";
    for (uint32_t k = 0; k < ni; ++k)
      if (!(scalar_mask & (1u << k))) o << "  vec_t a" << k << "[UNROLL];\n";
    o << "  #pragma unroll\n  for (int u = 0; u < UNROLL; ++u) {\n    const u64 idx = base + (u64)u * 256;\n    if (idx < nvec) {\n";
    for (uint32_t k = 0; k < ni; ++k)
      if (!(scalar_mask & (1u << k))) o << "      a" << k << "[u] = ldv(in" << k << " + idx * VEC);\n";
    o << "    }\n  }\n";
    o << "  #pragma unroll\n  for (int u = 0; u < UNROLL; ++u) {\n    const u64 idx = base + (u64)u * 256;\n    if (idx < nvec) {\n";
    for (uint32_t k = 0; k < no; ++k) o << "      vec_t r" << k << ";\n";
    o << "      #pragma unroll\n      for (int l = 0; l < VEC; ++l) {\n";
    for (uint32_t k = 0; k < ni; ++k) {
      if (scalar_mask & (1u << k)) o << "      const T v" << k << " = s" << k << ";\n";
      else o << "      const T v" << k << " = a" << k << "[u].x[l];\n";
    }
    o << body;
    for (uint32_t k = 0; k < no; ++k) o << "      r" << k << ".x[l] = " << prog.outputs[k] << ";\n";
    o << "      }\n";
    for (uint32_t k = 0; k < no; ++k) o << "      stv(out" << k << " + idx * VEC, r" << k << ");\n";
    o << "    }\n  }\n";
    // ragged tail (< VEC elements): first threads of block 0
    o << "  if (blockIdx.x == 0) {\n    const u64 g = nvec * VEC + threadIdx.x;\n    if (g < n) {\n";
    for (uint32_t k = 0; k < ni; ++k) {
      if (scalar_mask & (1u << k)) o << "      const T v" << k << " = s" << k << ";\n";
      else o << "      const T v" << k << " = in" << k << "[g];\n";
    }
    o << body;
    for (uint32_t k = 0; k < no; ++k) o << "      out" << k << "[g] = " << prog.outputs[k] << ";\n";
    o << "    }\n  }\n}\n";
  } else {
    // General MATLAB implicit expansion. Host coalesces dims (fused.cu) so rank <= 6 and strides are in
    // elements with 0 on broadcast dims — the same (shape,stride) encoding the wgpu provider uploads
    // (backend/wgpu/provider/ops/elementwise.rs:1669-1690), minus the 128-deep per-thread rank loop.
    o << "#define MAXR 6\n#define MAXIN " << (ni ? ni : 1) << "\n";
    o << "struct BParams { u64 len; u32 rank; u32 pad; u64 shape[MAXR]; u64 stride[MAXIN][MAXR]; };\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256) rm_fused_ew(" << input_params(ni);
    for (uint32_t k = 0; k < no; ++k) o << "T* __restrict__ out" << k << ", ";
    o << "const __grid_constant__ BParams p) {\n  RM_PDL_PROLOGUE();\n";
    o << "  for (u64 g = (u64)blockIdx.x * 256 + threadIdx.x; g < p.len; g += (u64)gridDim.x * 256) {\n";
    o << "    u64 rem = g;\n";
    for (uint32_t k = 0; k < ni; ++k) o << "    u64 i" << k << " = 0;\n";
    o << "    #pragma unroll\n    for (int d = 0; d < MAXR; ++d) {\n      if (d < (int)p.rank) {\n        const u64 dim = p.shape[d];\n        const u64 c = rem % dim; rem /= dim;\n";
    for (uint32_t k = 0; k < ni; ++k) o << "        i" << k << " += c * p.stride[" << k << "][d];\n";
    o << "      }\n    }\n    {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      const T v" << k << " = in" << k << "[i" << k << "];\n";
    o << body;
    for (uint32_t k = 0; k < no; ++k) o << "      out" << k << "[g] = " << prog.outputs[k] << ";\n";
    o << "    }\n  }\n}\n";
  }
  return o.str();
}

std::string emit_reduction_cuda(const ReductionProgram& prog, RedOp op, RedLayout layout) {
  std::ostringstream o;
  o << "#define RM_F32 " << (prog.scalar_ty == "f32" ? 1 : 0) << "\n#define RM_LIBM_TRIG " << libm_trig() << "\n" << kPrelude;
  const uint32_t ni = prog.n_inputs;
  const char* identity = op == RedOp::Sum ? "0.0" : op == RedOp::Prod ? "1.0" : op == RedOp::Max ? "-CUDART_INF" : "CUDART_INF";
  o << "#define CUDART_INF __longlong_as_double(0x7ff0000000000000LL)\n";
  o << "#define IDENT (" << identity << ")\n";
  o << "#define OMITNAN " << (prog.omit_nan ? 1 : 0) << "\n";
  // accumulate in f64 regardless of storage type (f32 storage: strictly more accurate than the shader's f32 tile)
  switch (op) {
    case RedOp::Sum: o << "#define COMBINE(a, b) ((a) + (b))\n#define TRACK_NAN 1\n"; break;
    case RedOp::Prod: o << "#define COMBINE(a, b) ((a) * (b))\n#define TRACK_NAN 1\n"; break;
    case RedOp::Max: o << "#define COMBINE(a, b) fmax((a), (b))\n#define TRACK_NAN 0\n"; break;   // f64::max ignores NaN (simple_provider.rs:7375)
    case RedOp::Min: o << "#define COMBINE(a, b) fmin((a), (b))\n#define TRACK_NAN 0\n"; break;
  }
  o << R"CUDA(
// canonical NaN, as the reference's reduction shader writes it (fusion.rs:1958-1962)
#define CANON_NAN __longlong_as_double(0x7ff8000000000000LL)
// NaN handling without per-element bookkeeping: in include mode a NaN term poisons the accumulator by itself (x + NaN,
// x * NaN) and finish() canonicalises; in omit mode NaN terms are replaced by the identity (branch-free select);
// max/min use fmax/fmin, which ignore NaN like f64::max (simple_provider.rs:7375).
__device__ __forceinline__ void accumulate(double& acc, T val) {
#if TRACK_NAN && OMITNAN
  acc = COMBINE(acc, (val != val) ? IDENT : (double)val);
#else
  acc = COMBINE(acc, (double)val);
#endif
}
__device__ __forceinline__ double warp_reduce(double v) {
  #pragma unroll
  for (int off = 16; off > 0; off >>= 1) { double o = __shfl_xor_sync(0xffffffffu, v, off); v = COMBINE(v, o); }
  return v;
}
// block-wide combine: warp shuffles, then one shared-memory hop, fixed order => deterministic
__device__ __forceinline__ double block_reduce(double v, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_reduce(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = IDENT;
  if (warp == 0) { r = lane < nwarp ? smem[lane] : IDENT; r = warp_reduce(r); }
  return r;  // valid in warp 0
}
__device__ __forceinline__ T finish(double acc, int use_div, double factor) {
  double r = use_div ? acc / factor : acc * factor;
  if (r != r) r = CANON_NAN;
  return (T)r;
}
// Fused exchange tail (comm.cu protocol): the block that produced the final scalar stores it into slot [bank][rank] of every
// rank's peer-mapped slot buffer and releases the step flag (system scope), one thread per destination rank. Peer stores go
// over NVLink; nothing waits here, the fold happens in the consumer's combine.
#define P2P_BANKS 8
#define P2P_MAXR 16
// Combine of an earlier exchange step fused in front of the publish (comm.cu p2p_fold_warp, same protocol): executed by warp 0 of
// the block that holds the final scalar. Lanes < n wait (bounded) for one rank's flag each in this rank's own slot buffer, the N
// values are folded in rank order, lane 0 writes the sum into the earlier step's result handle.
__device__ __forceinline__ void combine_prev(void* const* peers, u32 n, u32 rank, u64 prev_step1, T* prev_dst, int* err) {
  const u64 step = prev_step1 - 1;
  const char* base = (const char*)peers[rank];
  const u64 slot0 = (step % P2P_BANKS) * P2P_MAXR;
  const u32 q = threadIdx.x & 31;
  int bad = 0;
  double v = 0.0;
  if (q < n) {
    const unsigned long long* flag = (const unsigned long long*)(base + (u64)P2P_BANKS * P2P_MAXR * 8 + (slot0 + q) * 8);
    const long long t0 = clock64();
    for (;;) {
      unsigned long long f;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flag) : "memory");
      if (f >= step + 1) break;
      if (clock64() - t0 > 20000000000LL) { bad = 1; break; }
      __nanosleep(100);
    }
    v = *(volatile const double*)(base + (slot0 + q) * 8);
  }
  bad = __any_sync(0xffffffffu, bad);
  double s = 0.0;
  for (u32 r = 0; r < n; ++r) s += __shfl_sync(0xffffffffu, v, r);
  if (q == 0) {
    if (bad) atomicExch(err, 1);
    *prev_dst = (T)(bad ? CANON_NAN : s);
  }
}
__device__ __forceinline__ void publish_scalar(void* const* peers, u32 n, u32 rank, u64 step, double value) {
  if (threadIdx.x < n) {
    char* base = (char*)peers[threadIdx.x];
    const u64 slot = (step % P2P_BANKS) * P2P_MAXR + rank;
    *(double*)(base + slot * 8) = value;
    unsigned long long* flag = (unsigned long long*)(base + (u64)P2P_BANKS * P2P_MAXR * 8 + slot * 8);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(step + 1) : "memory");
  }
}
)CUDA";

  if (layout == RedLayout::Contig) {
    // slice s occupies [s*len, (s+1)*len). grid = (blocks_per_slice, num_slices).
    o << "#define RED_U " << red_unroll() << "\n";
    o << "#define RED_IDX(j) (" << (red_reverse() ? "(nvec - 1 - (j))" : "(j)") << ")\n";
    if (red_minblocks() > 0) o << "extern \"C\" __global__ void __launch_bounds__(256, " << red_minblocks() << ") rm_fused_red(";
    else o << "extern \"C\" __global__ void __launch_bounds__(256) rm_fused_red(";
    o << input_params(ni)
      << "T* __restrict__ out, double* __restrict__ partial, u32* __restrict__ pflags, u32* __restrict__ tickets,\n"
         "    u64 len, u64 num_slices, int vec_ok, int use_div, double factor, u32 bps, u64 inner, u32 sl,\n"
         "    void* const* __restrict__ pub_peers, u32 pub_n, u32 pub_rank, u64 pub_step, T* __restrict__ pub_prev_dst, u64 pub_prev_step1,\n    int* __restrict__ pub_err, double p0) {\n  RM_PDL_PROLOGUE();\n";
    o << "  __shared__ double smem[32];\n  __shared__ bool is_last;\n";
    o << "  const u64 slice = blockIdx.x / bps;\n  const u32 bidx = blockIdx.x % bps;\n  const u64 base = slice * len;\n";
    o << "  const u64 nvec = vec_ok ? len / VEC : 0;\n";
    o << "  const u64 tid = (u64)bidx * blockDim.x + threadIdx.x;\n  const u64 nthr = (u64)bps * blockDim.x;\n";
    // two independent accumulators (even / odd vector lanes): halves the dependent DADD chain, fixed fold order at the end
    o << "  double acc0 = IDENT, acc1 = IDENT;\n";
    // Work split. Grid-stride (thread t takes vectors t, t + nthr, ...) leaves the first (nvec mod nthr) threads one iteration more
    // than the rest (28 vs 27 on the headline reduction). The alternative -- every CTA owns one contiguous run of ceil(nvec / bps)
    // vectors -- balances that but measured SLOWER (r57: 46.1 us vs 44.3 us; sum(A) 24.4 vs 23.6): with grid-stride the resident
    // CTAs sweep one moving window of the array together (DRAM pages and L2 sets shared by neighbours), with blocked runs 592
    // separate streams are open at once. Grid-stride stays the default; RUNMAT_B200_RED_BLOCKED=1 selects the other.
    if (red_blocked()) {
      o << "  const u64 per_cta = (nvec + bps - 1) / bps;\n  const u64 lo = (u64)bidx * per_cta;\n"
           "  const u64 hi = lo + per_cta < nvec ? lo + per_cta : nvec;\n  const u64 vstep = blockDim.x;\n  u64 i = lo + threadIdx.x;\n";
    } else {
      o << "  const u64 hi = nvec;\n  const u64 vstep = nthr;\n  u64 i = tid;\n";
    }
    // U 256-bit vectors per input in flight per thread per iteration (all loads issued before any math)
    o << "  for (; i + (u64)(RED_U - 1) * vstep < hi; i += (u64)RED_U * vstep) {\n";
    for (uint32_t k = 0; k < ni; ++k) {
      o << "    vec_t a" << k << "[RED_U];\n";
    }
    o << "    #pragma unroll\n    for (int u = 0; u < RED_U; ++u) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      a" << k << "[u] = ldv(in" << k << " + base + RED_IDX(i + (u64)u * vstep) * VEC);\n";
    o << "    }\n";
    o << "    #pragma unroll\n    for (int u = 0; u < RED_U; ++u) {\n      #pragma unroll\n      for (int l = 0; l < VEC; ++l) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "        const T v" << k << " = a" << k << "[u].x[l];\n";
    o << "        if (l & 1) accumulate(acc1, " << prog.val_expr << "); else accumulate(acc0, " << prog.val_expr << ");\n      }\n    }\n  }\n";
    o << "  for (; i < hi; i += vstep) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "    const vec_t a" << k << " = ldv(in" << k << " + base + RED_IDX(i) * VEC);\n";
    o << "    #pragma unroll\n    for (int l = 0; l < VEC; ++l) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      const T v" << k << " = a" << k << ".x[l];\n";
    o << "      if (l & 1) accumulate(acc1, " << prog.val_expr << "); else accumulate(acc0, " << prog.val_expr << ");\n    }\n  }\n";
    o << "  for (u64 g = nvec * VEC + tid; g < len; g += nthr) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "    const T v" << k << " = in" << k << "[base + g];\n";
    o << "    accumulate(acc0, " << prog.val_expr << ");\n  }\n";
    o << R"CUDA(
  double total = block_reduce(COMBINE(acc0, acc1), smem);
  if (bps == 1) {
    if (threadIdx.x == 0) { const T r = finish(total, use_div, factor); out[slice] = r; smem[0] = (double)r; }
    if (pub_n) {
    __syncthreads();
    if (pub_prev_step1 && threadIdx.x < 32) combine_prev(pub_peers, pub_n, pub_rank, pub_prev_step1, pub_prev_dst, pub_err);
    publish_scalar(pub_peers, pub_n, pub_rank, pub_step, smem[0]);
  }
    return;
  }
  // two-stage, atomic-free on the data path: partials are combined by the LAST block in a fixed order
  if (threadIdx.x == 0) {
    partial[slice * bps + bidx] = total;
    __threadfence();
    const u32 t = atomicAdd(&tickets[slice], 1u);
    is_last = (t == bps - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc2 = IDENT;
  for (u32 b = threadIdx.x; b < bps; b += blockDim.x) acc2 = COMBINE(acc2, __ldcg(&partial[slice * bps + b]));
  double total2 = block_reduce(acc2, smem);
  if (threadIdx.x == 0) { const T r = finish(total2, use_div, factor); out[slice] = r; tickets[slice] = 0; smem[0] = (double)r; }
  if (pub_n) {
    __syncthreads();
    if (pub_prev_step1 && threadIdx.x < 32) combine_prev(pub_peers, pub_n, pub_rank, pub_prev_step1, pub_prev_dst, pub_err);
    publish_scalar(pub_peers, pub_n, pub_rank, pub_step, smem[0]);
  }
}
)CUDA";
  } else if (layout == RedLayout::Interleaved) {
    // Small power-of-two `inner` (e.g. the 8 images of mean(imgs,[2 3]) on a batch-fastest [B,H,W] tensor): block y of the
    // tensor, [inner x len], is one contiguous stream in which element e belongs to slice e % inner. Every thread sweeps
    // 256-bit vectors with a grid stride that is a multiple of inner, so vector lane l always lands on slice (b0 + l) % inner
    // and the VEC per-slice accumulators stay in registers; RED_IU vectors per input are in flight per iteration (the scalar
    // loads of the Strided kernel kept 16 B per thread in flight and ran at 0.41 of the copy bandwidth, r06). Lanes that
    // share a slice are folded with xor-shuffles, warps through shared memory in warp order, CTAs by the last CTA (ticket)
    // in CTA order: deterministic, no floating-point atomics. grid = (CTAs per block, blocks); R = max(inner, VEC).
    const int iu = ni <= 1 ? 4 : 2;
    o << "#define RED_IU " << iu << "\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256, 4) rm_fused_red(" << input_params(ni)
      << "T* __restrict__ out, double* __restrict__ partial, u32* __restrict__ pflags, u32* __restrict__ tickets,\n"
         "    u64 len, u64 num_slices, int vec_ok, int use_div, double factor, u32 bps, u64 inner, u32 sl,\n"
         "    void* const* __restrict__ pub_peers, u32 pub_n, u32 pub_rank, u64 pub_step, T* __restrict__ pub_prev_dst, u64 pub_prev_step1,\n    int* __restrict__ pub_err, double p0) {\n  RM_PDL_PROLOGUE();\n";
    o << "  __shared__ double shw[8 * 256];\n  __shared__ bool is_last;\n";
    o << "  const u32 in_ = (u32)inner, R = in_ > VEC ? in_ : VEC, G = R / VEC;\n";
    o << "  const u64 blk = inner * len, pbase = (u64)blockIdx.y * blk;\n";
    o << "  const u64 nvec = blk / VEC, nthr = (u64)gridDim.x * 256;\n";
    o << "  const u64 v0 = (u64)blockIdx.x * 256 + threadIdx.x;\n";
    o << "  const u32 b0 = (u32)((v0 * VEC) % R);\n";
    o << "  double acc[VEC];\n  #pragma unroll\n  for (int l = 0; l < VEC; ++l) acc[l] = IDENT;\n";
    o << "  u64 v = v0;\n";
    o << "  for (; v + (u64)(RED_IU - 1) * nthr < nvec; v += (u64)RED_IU * nthr) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "    vec_t a" << k << "[RED_IU];\n";
    o << "    #pragma unroll\n    for (int u = 0; u < RED_IU; ++u) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      a" << k << "[u] = ldv(in" << k << " + pbase + (v + (u64)u * nthr) * VEC);\n";
    o << "    }\n    #pragma unroll\n    for (int u = 0; u < RED_IU; ++u) {\n      #pragma unroll\n      for (int l = 0; l < VEC; ++l) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "        const T v" << k << " = a" << k << "[u].x[l];\n";
    o << "        accumulate(acc[l], " << prog.val_expr << ");\n      }\n    }\n  }\n";
    o << "  for (; v < nvec; v += nthr) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "    const vec_t a" << k << " = ldv(in" << k << " + pbase + v * VEC);\n";
    o << "    #pragma unroll\n    for (int l = 0; l < VEC; ++l) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      const T v" << k << " = a" << k << ".x[l];\n";
    o << "      accumulate(acc[l], " << prog.val_expr << ");\n    }\n  }\n";
    o << R"CUDA(
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lanes q and q' of a warp hold the same slices iff q % G == q' % G (G is a power of two <= 32)
  #pragma unroll
  for (int l = 0; l < VEC; ++l)
    for (u32 off = 16; off >= G && off > 0; off >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, acc[l], off); acc[l] = COMBINE(acc[l], ov); }
  if (lane < G) {
    #pragma unroll
    for (int l = 0; l < VEC; ++l) shw[warp * R + b0 + l] = acc[l];
  }
  __syncthreads();
  // slot s of a warp row belongs to slice s % inner (more than one slot per slice only when inner < VEC)
  double t = IDENT;
  if (threadIdx.x < in_) {
    for (u32 w = 0; w < 8; ++w)
      for (u32 s = threadIdx.x; s < R; s += in_) t = COMBINE(t, shw[w * R + s]);
  }
  if (gridDim.x == 1) {
    if (threadIdx.x < in_) out[(u64)blockIdx.y * in_ + threadIdx.x] = finish(t, use_div, factor);
    return;
  }
  if (threadIdx.x < in_) partial[((u64)blockIdx.y * gridDim.x + blockIdx.x) * in_ + threadIdx.x] = t;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { const u32 tk = atomicAdd(&tickets[blockIdx.y], 1u); is_last = (tk == gridDim.x - 1); }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    // last CTA of the block: thread q walks CTA partials q / inner, q / inner + W, ... of slice q % inner in ascending order,
    // then the W thread partials of a slice are folded in thread order
    const u32 W = 256 / in_;
    const u32 fs = threadIdx.x % in_, fy = threadIdx.x / in_;
    double acc2 = IDENT;
    if (fy < W)
      for (u32 y = fy; y < gridDim.x; y += W) acc2 = COMBINE(acc2, __ldcg(&partial[((u64)blockIdx.y * gridDim.x + y) * in_ + fs]));
    shw[threadIdx.x] = acc2;   // every read of shw above happened before the two barriers
    __syncthreads();
    if (threadIdx.x < in_) {
      for (u32 w = 1; w < W; ++w) acc2 = COMBINE(acc2, shw[w * in_ + fs]);
      out[(u64)blockIdx.y * in_ + fs] = finish(acc2, use_div, factor);
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.y] = 0;
}
)CUDA";
  } else {
    // Strided layout: element r of slice s lives at sbase(s) + r*inner. A CTA owns `sl` adjacent slices x (256/sl) row
    // lanes, so a warp always touches contiguous memory even when there are only a handful of slices (e.g. the 8 images
    // of mean(imgs,[2 3])); lanes are combined through shared memory in a fixed order, chunks (gridDim.y) through the
    // same deterministic last-block finish as the contiguous layout.
    o << "extern \"C\" __global__ void __launch_bounds__(256) rm_fused_red(" << input_params(ni)
      << "T* __restrict__ out, double* __restrict__ partial, u32* __restrict__ pflags, u32* __restrict__ tickets,\n"
         "    u64 len, u64 num_slices, int vec_ok, int use_div, double factor, u32 bps, u64 inner, u32 sl,\n"
         "    void* const* __restrict__ pub_peers, u32 pub_n, u32 pub_rank, u64 pub_step, T* __restrict__ pub_prev_dst, u64 pub_prev_step1,\n    int* __restrict__ pub_err, double p0) {\n  RM_PDL_PROLOGUE();\n";
    o << "  __shared__ bool is_last;\n  __shared__ double sacc[256];\n";
    o << "  const u32 sloc = threadIdx.x % sl, lane = threadIdx.x / sl, rl = blockDim.x / sl;\n";
    o << "  const u64 s = (u64)blockIdx.x * sl + sloc;\n";
    o << "  const u64 chunk = (len + gridDim.y - 1) / gridDim.y;\n";
    o << "  const u64 r0 = (u64)blockIdx.y * chunk; const u64 r1 = r0 + chunk < len ? r0 + chunk : len;\n";
    o << "  double acc = IDENT, accb = IDENT;\n";
    o << "  const u64 sbase = (s % inner) + (s / inner) * inner * len;\n";
    o << "  if (s < num_slices) {\n    u64 r = r0 + lane;\n";
    o << "    for (; r + 3 * (u64)rl < r1; r += 4 * (u64)rl) {\n";
    for (int u = 0; u < 4; ++u)
      for (uint32_t k = 0; k < ni; ++k) o << "      const T w" << u << "_" << k << " = in" << k << "[sbase + (r + " << u << " * (u64)rl) * inner];\n";
    for (int u = 0; u < 4; ++u) {
      o << "      {\n";
      for (uint32_t k = 0; k < ni; ++k) o << "        const T v" << k << " = w" << u << "_" << k << ";\n";
      o << "        accumulate(" << ((u & 1) ? "accb" : "acc") << ", " << prog.val_expr << ");\n      }\n";
    }
    o << "    }\n    for (; r < r1; r += rl) {\n";
    for (uint32_t k = 0; k < ni; ++k) o << "      const T v" << k << " = in" << k << "[sbase + r * inner];\n";
    o << "      accumulate(acc, " << prog.val_expr << ");\n    }\n  }\n";
    o << R"CUDA(
  acc = COMBINE(acc, accb);
  sacc[threadIdx.x] = acc;
  __syncthreads();
  if (lane == 0) {
    for (u32 l = 1; l < rl; ++l) acc = COMBINE(acc, sacc[l * sl + sloc]);
  }
  const bool owner = lane == 0 && s < num_slices;
  if (gridDim.y == 1) {
    if (owner) out[s] = finish(acc, use_div, factor);
    return;
  }
  if (owner) partial[(u64)blockIdx.y * num_slices + s] = acc;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { const u32 t = atomicAdd(&tickets[blockIdx.x], 1u); is_last = (t == gridDim.y - 1); }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // Last CTA: all 256 threads fold the gridDim.y chunk partials of this CTA's sl slices (a handful of owner threads walking
  // ~1000 partials each was a ~200 us serial tail on mean(imgs,[2 3])). Thread t walks chunks t/sl, t/sl + W, ... of slice
  // t % sl in ascending order; the W thread partials of a slice are then folded in thread order: fixed order, deterministic.
  {
    const u32 W = 256 / sl;
    const u32 fs = threadIdx.x % sl, fy = threadIdx.x / sl;
    const u64 gs = (u64)blockIdx.x * sl + fs;
    double acc2 = IDENT;
    if (gs < num_slices)
      for (u32 y = fy; y < gridDim.y; y += W) acc2 = COMBINE(acc2, __ldcg(&partial[(u64)y * num_slices + gs]));
    __syncthreads();  // every thread has finished reading sacc from the first fold
    sacc[threadIdx.x] = acc2;
    __syncthreads();
    if (threadIdx.x < sl && gs < num_slices) {
      for (u32 w = 1; w < W; ++w) acc2 = COMBINE(acc2, sacc[w * sl + fs]);
      out[gs] = finish(acc2, use_div, factor);
    }
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;
}
)CUDA";
  }
  return o.str();
}

}  // namespace rm
