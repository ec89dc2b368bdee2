// Host check of the generated kernels' f64 sin / cos (the "rm_trig" block of fusion_lower.cpp's prelude). Test infrastructure:
// tests/test_lowering.py::test_lean_trig_against_libm lowers a program, cuts the block out of the generated CUDA source into
// fused_trig_block.inc, compiles this file with g++ -ffp-contract=off and runs it. The reference is glibc in long double.
// Prints the maxima in ulp; exit 1 if any exceeds the bound given on the command line (default 3 ulp), or a special value differs.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <random>
typedef unsigned int u32;
#define __constant__ static const
#define __device__
#define __noinline__
#define __forceinline__ inline
static inline int __double2hiint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(b >> 32); }
static inline int __double2loint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo) { uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d; }
#include "fused_trig_block.inc"

static double ulps(double a, long double ref) {
  if (ref == 0) return a == 0 ? 0 : 1e300;
  int e;
  frexpl(ref, &e);
  return (double)(fabsl((long double)a - ref) / ldexpl(1.0L, e - 53));
}

int main(int argc, char** argv) {
  const double bound = argc > 1 ? atof(argv[1]) : 3.0;
  std::mt19937_64 g(1);
  double ws = 0, wc = 0, xs = 0, xc = 0;
  auto test = [&](double x) {
    const double us = ulps(rm_sin(x), sinl((long double)x)), uc = ulps(rm_cos(x), cosl((long double)x));
    if (us > ws) { ws = us; xs = x; }
    if (uc > wc) { wc = uc; xc = x; }
  };
  std::uniform_real_distribution<double> U(0, 4 * M_PI), V(-105615, 105615), W(-40, 40);
  for (int i = 0; i < 2000000; ++i) { test(U(g)); test(V(g)); test(W(g)); }
  for (int k = -33619; k <= 33619; ++k)        // the doubles next to every multiple of pi/2 of the fast range (zeros of sin and cos)
    for (int d = -2; d <= 2; ++d) {
      double y = k * M_PI_2;
      for (int j = 0; j < abs(d); ++j) y = nextafter(y, d > 0 ? 1e300 : -1e300);
      if (fabs(y) < 105615.0) test(y);
    }
  for (int e = -40; e < 3; ++e)
    for (int i = 0; i < 500; ++i) { const double x = ldexp(1.0 + (g() % 1000000) / 1e6, e); test(x); test(-x); }
  printf("max ulp: sin %.3f at %.17g, cos %.3f at %.17g\n", ws, xs, wc, xc);
  int bad = ws > bound || wc > bound;
  // special values: the fdlibm / libm answers
  const double s0 = rm_sin(-0.0);
  bad |= !(s0 == 0.0 && std::signbit(s0)) || std::signbit(rm_sin(0.0)) || rm_cos(0.0) != 1.0 || rm_cos(-0.0) != 1.0;
  bad |= rm_sin(1e-300) != 1e-300 || rm_sin(-5e-324) != -5e-324 || rm_cos(1e-30) != 1.0;
  bad |= rm_sin(1e6) != sin(1e6) || rm_cos(-1e22) != cos(-1e22) || rm_sin(105615.0) != sin(105615.0);   // out-of-line library route
  bad |= !std::isnan(rm_sin(INFINITY)) || !std::isnan(rm_cos(-INFINITY)) || !std::isnan(rm_sin(NAN)) || !std::isnan(rm_cos(NAN));
  // the reference's own GPU round-trip test compares sin([0 1 2 3]) with the host's f64::sin by assert_eq (sin.rs:620-634,
  // sin_gpu_provider_roundtrip): these four must be bit-equal to libm, not just close
  for (double x : {0.0, 1.0, 2.0, 3.0}) bad |= rm_sin(x) != sin(x);
  if (bad) printf("FAILED\n");
  return bad;
}
