// Host check of runmat_b200/csrc/mc_math.h against glibc's libm (what Rust's f64::ln / exp / sin / cos resolve to on Linux):
// maximum relative error of neg2log_u53, exp_fast and sincos_turn_u53 over the argument ranges the Monte-Carlo kernel produces.
//   g++ -O2 -ffp-contract=off -o check_mc_math check_mc_math.cpp && ./check_mc_math      (prints the maxima; exit 1 above 2e-15)
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include "../../runmat_b200/csrc/mc_math.h"
int main() {
  uint64_t s = 0x9e3779b97f4a7c15ull;
  auto u53 = [&]() { s = s * 6364136223846793005ull + 1; return s >> 11; };
  const double inv = 1.0 / 9007199254740992.0;
  const double* tab = rm_mc::kNeg2LogTab;
  double elog = 0, eexp = 0, esc = 0;
  auto check_log = [&](uint64_t x) {
    double u = (double)x * inv;
    if (u <= 0.0) u = 2.2250738585072014e-308;
    const double w = -2.0 * log(u), g = rm_mc::neg2log_u53(x, tab);
    elog = fmax(elog, fabs(g - w) / fabs(w));
  };
  auto check_sc = [&](uint64_t x) {
    double sn, cs;
    rm_mc::sincos_turn_u53(x, &sn, &cs);
    const double ang = 2.0 * 3.14159265358979323846 * ((double)x * inv);
    esc = fmax(esc, fmax(fabs(sn - sin(ang)), fabs(cs - cos(ang))));  // absolute: the host's own angle rounding is ~4e-16
  };
  for (int i = 0; i < 4000000; ++i) {
    uint64_t x = u53();
    if (i % 7 == 0) x >>= (i % 53);                                   // small uniforms: large |log|
    if (i % 11 == 0) x = (1ull << 53) - 1 - (u53() >> (i % 53));       // just below one: tiny |log|
    if (i % 13 == 0) x = ((1ull << (1 + i % 52)) - 1) - (i % 3);       // just below a power of two
    if (x < (1ull << 53)) check_log(x);
    const double v = ((double)u53() * inv - 0.5) * (i % 3 == 0 ? 1400.0 : (i % 3 == 1 ? 20.0 : 0.2));
    if (fabs(v) < 700) { const double w = exp(v), g = rm_mc::exp_fast(v); eexp = fmax(eexp, fabs(g - w) / w); }
    check_sc(u53());
  }
  // every table bucket edge, both sides, at several exponents
  for (int e = 0; e < 53; e += 4)
    for (int i = 0; i <= 256; ++i)
      for (int d = -2; d <= 2; ++d) {
        const uint64_t x = (((1ull << 52) + ((uint64_t)i << 44)) + (uint64_t)(int64_t)d) >> e;
        if (x > 0 && x < (1ull << 53)) check_log(x);
      }
  for (uint64_t x : {0ull, 1ull, 2ull, 3ull, (1ull << 53) - 1, (1ull << 52), (1ull << 52) - 1, (1ull << 52) + 1}) check_log(x);
  for (uint64_t q = 0; q < 8; ++q)
    for (int d = -2; d <= 2; ++d) { const uint64_t x = (q << 50) + (uint64_t)(int64_t)d; if (x < (1ull << 53)) check_sc(x); }
  check_sc((1ull << 53) - 1);
  printf("max rel err: neg2log_u53 %.3e  exp_fast %.3e   max abs err sincos_turn_u53 %.3e   -2 log(MIN_POSITIVE) %.17g vs %.17g\n", elog, eexp, esc,
         rm_mc::neg2log_u53(0, tab), -2.0 * log(2.2250738585072014e-308));
  printf("exp_fast(800) %g exp_fast(-800) %g exp_fast(nan) %g\n", rm_mc::exp_fast(800.0), rm_mc::exp_fast(-800.0), rm_mc::exp_fast(NAN));
  return (elog < 2e-15 && eexp < 2e-15 && esc < 2e-15) ? 0 : 1;
}
