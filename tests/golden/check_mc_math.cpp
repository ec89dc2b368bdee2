// Host check of runmat_b200/csrc/mc_math.h against glibc's libm (what Rust's f64::ln / exp / sin / cos resolve to on Linux):
// maximum relative error of log_unit, exp_fast and sincos_turn over the argument ranges the Monte-Carlo kernel produces.
//   g++ -O2 -ffp-contract=off -o check_mc_math check_mc_math.cpp && ./check_mc_math      (prints the maxima; exit 1 above 2e-15)
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include "../../runmat_b200/csrc/mc_math.h"
int main() {
  uint64_t s = 0x9e3779b97f4a7c15ull;
  auto uni = [&]() { s = s * 6364136223846793005ull + 1; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
  double elog = 0, eexp = 0, esc = 0;
  for (int i = 0; i < 4000000; ++i) {
    double u = uni();
    if (i % 7 == 0) u = ldexp(u, -(i % 1000));          // small uniforms: large |log|
    if (i % 11 == 0) u = 1.0 - ldexp(uni(), -(i % 50));  // just below one: tiny |log|
    if (u <= 0.0) u = 2.2250738585072014e-308;
    if (u < 1.0) { const double w = log(u), g = rm_mc::log_unit(u); elog = fmax(elog, fabs(g - w) / fabs(w)); }
    const double x = (uni() - 0.5) * (i % 3 == 0 ? 1400.0 : (i % 3 == 1 ? 20.0 : 0.2));
    if (fabs(x) < 700) { const double w = exp(x), g = rm_mc::exp_fast(x); eexp = fmax(eexp, fabs(g - w) / w); }
    const double v = uni();
    double sn, cs;
    rm_mc::sincos_turn(v, &sn, &cs);
    const double ang = 2.0 * 3.14159265358979323846 * v;
    esc = fmax(esc, fmax(fabs(sn - sin(ang)), fabs(cs - cos(ang))));  // absolute: the host's own angle rounding is ~4e-16
  }
  double sn, cs;
  for (double v : {0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 0.9999999999999999}) {
    rm_mc::sincos_turn(v, &sn, &cs);
    const double ang = 2.0 * 3.14159265358979323846 * v;
    esc = fmax(esc, fmax(fabs(sn - sin(ang)), fabs(cs - cos(ang))));
  }
  const double lmin = rm_mc::log_unit(2.2250738585072014e-308);
  printf("max rel err: log_unit %.3e  exp_fast %.3e   max abs err sincos_turn %.3e   log(MIN_POSITIVE) %.17g vs %.17g\n", elog, eexp, esc, lmin, log(2.2250738585072014e-308));
  printf("exp_fast(800) %g exp_fast(-800) %g exp_fast(nan) %g\n", rm_mc::exp_fast(800.0), rm_mc::exp_fast(-800.0), rm_mc::exp_fast(NAN));
  return (elog < 2e-15 && eexp < 2e-15 && esc < 2e-15) ? 0 : 1;
}
