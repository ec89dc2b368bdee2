// mc_math.h — lean f64 log / exp / sincos(2*pi*u) for the Monte-Carlo and Box-Muller kernels (mc.cu).
//
// Why not the CUDA math library here: r06 SASS of evolve_kernel: ~350 instructions per Box-Muller pair-step of which only 117 are
// FP64 arithmetic; ~110 are UMOV / IMAD.MOV that rebuild the library's polynomial coefficients as immediates every iteration
// (ptxas keeps the kernel at 32 registers and rematerialises instead of hoisting), and the kernel is issue-bound (SM 84 % busy,
// FP64 pipe 60 %). The versions below keep every coefficient in __constant__ memory, so each Horner step is ONE DFMA with a
// constant-bank operand, and use argument ranges this kernel actually has (u in (0,1) for log, a turn fraction for sincos).
// Accuracy: Taylor / atanh series truncated below 2e-16 relative, arguments reduced exactly or with a hi/lo split: a few ulp,
// far inside the 1e-10 parity bar against the host's libm (tests/test_gpu_parity.py: stochastic_evolution, random_normal).
// The same code compiles for the host (tests/golden/check_mc_math.cpp) where it is compared with glibc.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define RM_MC_HD __host__ __device__ __forceinline__
#define RM_MC_CONST __constant__
#else
#define RM_MC_HD inline
#define RM_MC_CONST static const
#endif

namespace rm_mc {

// 1/k!  (exp on |r| <= ln2/2: 0.3466^13/13! = 1.7e-16)
RM_MC_CONST double kExp[13] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800,
                               1.0 / 39916800, 1.0 / 479001600};
// 1/(2i+1)  (log m = 2 atanh(s), |s| <= 0.1716: s^20/21 < 2e-17 relative to s)
RM_MC_CONST double kAtanh[10] = {1.0, 1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19};
// sin x = x * sum (-1)^i x^(2i)/(2i+1)!, cos x = sum (-1)^i x^(2i)/(2i)!   on |x| <= pi/4
RM_MC_CONST double kSin[8] = {1.0, -1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800.0, -1.0 / 1307674368000.0};
RM_MC_CONST double kCos[9] = {1.0, -0.5, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600, -1.0 / 87178291200.0, 1.0 / 20922789888000.0};

RM_MC_HD double bits_to_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
RM_MC_HD uint64_t double_to_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

// exp(x). |x| < 700 takes the polynomial path (result is a normal number); everything else (overflow, underflow, NaN) goes to libm.
RM_MC_HD double exp_fast(double x) {
  if (!(fabs(x) < 700.0)) return exp(x);
  const double k = rint(x * 1.4426950408889634074);
  double r = fma(k, -6.93147180369123816490e-01, x);   // ln2 hi (trailing zeros: k * hi is exact)
  r = fma(k, -1.90821492927058770002e-10, r);          // ln2 lo
  // two independent Horner chains (even / odd powers) in r^2: half the dependent-DFMA depth of a plain degree-12 Horner (r24 ncu:
  // FP64 pipe 67 % and issue 70 % busy -- the kernel waits on dependent chains, not on a pipe)
  const double r2 = r * r;
  double pe = kExp[12], po = kExp[11];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 10; i >= 0; i -= 2) { pe = fma(pe, r2, kExp[i]); if (i >= 1) po = fma(po, r2, kExp[i - 1]); }
  const double p = fma(po, r, pe);
  return p * bits_to_double((uint64_t)((int64_t)k + 1023) << 52);  // exact scaling by 2^k, -1010 <= k <= 1010
}

// log(u) for a normal positive u <= 1 (the Box-Muller uniform, clamped to f64::MIN_POSITIVE by the caller).
RM_MC_HD double log_unit(double u) {
  const uint64_t b = double_to_bits(u);
  int e = (int)(b >> 52) - 1023;
  double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);  // [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }                                 // [0.7071, 1.4142]
  const double s = (m - 1.0) / (m + 1.0);
  const double s2 = s * s;
  const double s4 = s2 * s2;
  double pa = kAtanh[8], pb = kAtanh[9];  // even / odd coefficients of the series in s2: two chains in s4
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 6; i >= 0; i -= 2) { pa = fma(pa, s4, kAtanh[i]); pb = fma(pb, s4, kAtanh[i + 1]); }
  const double p = fma(pb, s2, pa);
  const double lm = 2.0 * s * p;  // log(m)
  const double ed = (double)e;
  return fma(ed, 6.93147180369123816490e-01, fma(ed, 1.90821492927058770002e-10, lm));
}

// sin and cos of 2*pi*u for u in [0, 1): quadrant by exact arithmetic on the turn fraction, Taylor polynomials on |x| <= pi/4.
RM_MC_HD void sincos_turn(double u, double* sn, double* cs) {
  const double t = 4.0 * u;       // quarter turns, exact
  const double q = rint(t);       // 0 .. 4
  const double x = (t - q) * 1.5707963267948966192;  // (t - q) is exact, |x| <= pi/4
  const double x2 = x * x;
  double ps = kSin[7], pc = kCos[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 6; i >= 0; --i) ps = fma(ps, x2, kSin[i]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 7; i >= 0; --i) pc = fma(pc, x2, kCos[i]);
  const double sx = x * ps, cx = pc;
  const int qi = (int)q & 3;
  const double s0 = (qi & 1) ? cx : sx, c0 = (qi & 1) ? sx : cx;
  *sn = (qi & 2) ? -s0 : s0;                 // quadrants: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s)
  *cs = (qi == 1 || qi == 2) ? -c0 : c0;
}

}  // namespace rm_mc
