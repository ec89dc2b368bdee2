// mc_math.h — lean f64 exp / -2 log(u) / sincos(2*pi*u) for the Monte-Carlo and Box-Muller kernels (mc.cu).
//
// Why not the CUDA math library here: r06 SASS of evolve_kernel: ~350 instructions per Box-Muller pair-step of which only 117 are
// FP64 arithmetic; ~110 are UMOV / IMAD.MOV that rebuild the library's polynomial coefficients as immediates every iteration
// (ptxas keeps the kernel at 32 registers and rematerialises instead of hoisting), and the kernel is issue-bound (SM 84 % busy,
// FP64 pipe 60 %). The versions below keep every coefficient in __constant__ memory, so each Horner step is ONE DFMA with a
// constant-bank operand, and use argument ranges this kernel actually has (u in (0,1) for log, a turn fraction for sincos).
// Accuracy: series / minimax polynomials truncated below 2e-16 relative, arguments reduced exactly or with a hi/lo split: a few
// ulp, far inside the 1e-10 parity bar against the host's libm (tests/test_gpu_parity.py: stochastic_evolution, random_normal).
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

// ---- Box-Muller pieces taken straight from the 53-bit LCG integers (x = state >> 11, u = x * 2^-53) ---------------------------
// r36 ncu/SASS of evolve_kernel: 160 instructions per pair-step of which ~70 go to the FP64 pipe (2 issue cycles each on B200):
// the uniform conversions (I2F.F64.U64 + DMUL), the division inside the atanh form of log (MUFU.RCP64H + 7 DFMA), a 10-term
// series, rint/F2I for the quadrant. The versions below start from the INTEGER: exponent and mantissa of u by count-leading-zeros
// and shifts (ALU pipe), a 256-entry table {1/c, -2 log c} (shared memory on the device) that leaves |r| = |m/c - 1| <= 2^-8 for
// a degree-7 log1p, the quadrant and the reduced angle by integer arithmetic, and fdlibm's minimax kernels for sin / cos on
// |x| <= pi/4 (k_sin.c / k_cos.c coefficients, error < 2^-58). ~37 FP64-pipe instructions per pair-step.

// {rc_i, -2 * log(c_i')} for mantissa bucket i = top 8 fraction bits; rc_i = fl(1 / (1 + (i + 0.5) / 256)); the second entry is
// computed from the ROUNDED rc_i (50-digit arithmetic, scripts in the commit message), with the ln 2 of the re-centring
// (buckets >= 106, i.e. m >= 1.414: m/2 and e + 1) folded in. Buckets 0 and 255 use c' = 1 exactly, so log(u) for u next to 1
// (and next to any power of two from below) keeps full RELATIVE accuracy.
#if defined(__CUDACC__)
#define RM_MC_TABLE __device__ const
#else
#define RM_MC_TABLE static const
#endif
RM_MC_TABLE double kNeg2LogTab[512] = {
  0x1.0000000000000p+0, 0.0,    0x1.fd04794a10e6ap-1, -0x1.7ee11ebd82ec4p-7,
  0x1.fb0c610d5e939p-1, -0x1.3e7295d25a7d5p-6,    0x1.f9182b6813bafp-1, -0x1.bcf712c743853p-6,
  0x1.f727cce5f530ap-1, -0x1.1d7f7eb9eebf1p-5,    0x1.f53b3a3fa204ep-1, -0x1.5c45a51b8d393p-5,
  0x1.f3526859b8cecp-1, -0x1.9ace7551cc515p-5,    0x1.f16d4c4401f17p-1, -0x1.d91a66c543cbep-5,
  0x1.ef8bdb389ebadp-1, -0x1.0b94f7c196173p-4,    0x1.edae0a9b3d3a5p-1, -0x1.2a7ec2214e879p-4,
  0x1.ebd3cff850b0cp-1, -0x1.494acc34d911dp-4,    0x1.e9fd21044e799p-1, -0x1.67f94f094bd92p-4,
  0x1.e829f39aef509p-1, -0x1.868a83083f6d0p-4,    0x1.e65a3dbe74d6bp-1, -0x1.a4fe9ffa3d233p-4,
  0x1.e48df596f3394p-1, -0x1.c355dd0921f2fp-4,    0x1.e2c511719ee16p-1, -0x1.e19070c276010p-4,
  0x1.e0ff87c01e100p-1, -0x1.ffae9119b92fbp-4,    0x1.df3d4f17de4dbp-1, -0x1.0ed839b5526fep-3,
  0x1.dd7e5e316d94cp-1, -0x1.1dcb263db1944p-3,    0x1.dbc2abe7d71d4p-1, -0x1.2cb0283f5de22p-3,
  0x1.da0a2f3803b41p-1, -0x1.3b87598b1b6f0p-3,    0x1.d854df401d855p-1, -0x1.4a50d3aa1b03fp-3,
  0x1.d6a2b33ef7448p-1, -0x1.590cafdf01c26p-3,    0x1.d4f3a293769cap-1, -0x1.67bb0726ec0fbp-3,
  0x1.d347a4bc01d34p-1, -0x1.765bf23a6be17p-3,    0x1.d19eb155f08a4p-1, -0x1.84ef898e82828p-3,
  0x1.cff8c01cff8c0p-1, -0x1.9375e55595edfp-3,    0x1.ce55c8eac7900p-1, -0x1.a1ef1d8061cd8p-3,
  0x1.ccb5c3b636e3ap-1, -0x1.b05b49bee4403p-3,    0x1.cb18a8930de60p-1, -0x1.beba818146764p-3,
  0x1.c97e6fb15e44dp-1, -0x1.cd0cdbf8c13e0p-3,    0x1.c7e7115d0ce95p-1, -0x1.db5270187d925p-3,
  0x1.c65285fd56843p-1, -0x1.e98b54967146bp-3,    0x1.c4c0c61456a8ep-1, -0x1.f7b79fec37de2p-3,
  0x1.c331ca3e91679p-1, -0x1.02ebb42bf3d4ap-2,    0x1.c1a58b327f576p-1, -0x1.09f561ee719c4p-2,
  0x1.c01c01c01c01cp-1, -0x1.10f8e422539b1p-2,    0x1.be9526d0769fap-1, -0x1.17f6458fca611p-2,
  0x1.bd10f365451b6p-1, -0x1.1eed90e2dc2c3p-2,    0x1.bb8f609879493p-1, -0x1.25ded0abc6ad3p-2,
  0x1.ba10679bd8488p-1, -0x1.2cca0f5f5f252p-2,    0x1.b89401b89401cp-1, -0x1.33af575770e4dp-2,
  0x1.b71a284ee6b34p-1, -0x1.3a8eb2d31a375p-2,    0x1.b5a2d4d5b081fp-1, -0x1.41682bf727bbfp-2,
  0x1.b42e00da17007p-1, -0x1.483bccce6e3dcp-2,    0x1.b2bba5ff26a23p-1, -0x1.4f099f4a230b1p-2,
  0x1.b14bbdfd760e6p-1, -0x1.55d1ad4232d70p-2,    0x1.afde42a2cb482p-1, -0x1.5c940075972b9p-2,
  0x1.ae732dd1c2a09p-1, -0x1.6350a28aaa759p-2,    0x1.ad0a798177693p-1, -0x1.6a079d0f7aad0p-2,
  0x1.aba41fbd2e5b1p-1, -0x1.70b8f97a1aa74p-2,    0x1.aa401aa401aa4p-1, -0x1.7764c128f2127p-2,
  0x1.a8de64688ebabp-1, -0x1.7e0afd630c276p-2,    0x1.a77ef750a56dap-1, -0x1.84abb75865137p-2,
  0x1.a621cdb4f8fdfp-1, -0x1.8b46f8223625bp-2,    0x1.a4c6e200d2637p-1, -0x1.91dcc8c340bdfp-2,
  0x1.a36e2eb1c432dp-1, -0x1.986d3228180c8p-2,    0x1.a217ae575ff2fp-1, -0x1.9ef83d2769a34p-2,
  0x1.a0c35b92ecdf1p-1, -0x1.a57df28244dcbp-2,    0x1.9f713117200d0p-1, -0x1.abfe5ae46124ap-2,
  0x1.9e2129a7d5f0ap-1, -0x1.b2797ee46320cp-2,    0x1.9cd34019cd340p-1, -0x1.b8ef670420c3bp-2,
  0x1.9b876f5262dd1p-1, -0x1.bf601bb0e44e0p-2,    0x1.9a3db2474fb98p-1, -0x1.c5cba543ae424p-2,
  0x1.98f603fe670a0p-1, -0x1.cc320c0176501p-2,    0x1.97b05f8d56652p-1, -0x1.d293581b6b3e7p-2,
  0x1.966cc01966cc0p-1, -0x1.d8ef91af31d5ep-2,    0x1.952b20d73ee97p-1, -0x1.df46c0c722d30p-2,
  0x1.93eb7d0aa6759p-1, -0x1.e598ed5a87e2ep-2,    0x1.92add0064ab74p-1, -0x1.ebe61f4dd7b0bp-2,
  0x1.9172152b841ddp-1, -0x1.f22e5e72f105cp-2,    0x1.903847ea1cec1p-1, -0x1.f871b28955045p-2,
  0x1.8f0063c018f00p-1, -0x1.feb0233e607cep-2,    0x1.8dca64397e408p-1, -0x1.0274dc16c232fp-1,
  0x1.8c9644f01efbcp-1, -0x1.058f3c703ebc5p-1,    0x1.8b64018b64019p-1, -0x1.08a73667c57aep-1,
  0x1.8a3395c018a34p-1, -0x1.0bbccdb0d24bcp-1,    0x1.8904fd503744bp-1, -0x1.0ed005f657da5p-1,
  0x1.87d8340ab6e97p-1, -0x1.11e0e2dad9cb6p-1,    0x1.86ad35cb59a84p-1, -0x1.14ef67f88685ap-1,
  0x1.8583fe7a7c018p-1, -0x1.17fb98e15095ep-1,    0x1.845c8a0ce5129p-1, -0x1.1b05791f07b4ap-1,
  0x1.8336d48397a24p-1, -0x1.1e0d0c33716bdp-1,    0x1.8212d9eba4018p-1, -0x1.211255986160cp-1,
  0x1.80f0965dfabcbp-1, -0x1.241558bfd1405p-1,    0x1.7fd005ff40180p-1, -0x1.27161913f853dp-1,
  0x1.7eb124ffa053bp-1, -0x1.2a1499f762bcap-1,    0x1.7d93ef9aa4b46p-1, -0x1.2d10dec508582p-1,
  0x1.7c7862170949fp-1, -0x1.300aead06350cp-1,    0x1.7b5e78c693733p-1, -0x1.3302c1658658ap-1,
  0x1.7a463005e918cp-1, -0x1.35f865c93293ep-1,    0x1.792f843c689c3p-1, -0x1.38ebdb38ed320p-1,
  0x1.781a71dc01782p-1, -0x1.3bdd24eb14b69p-1,    0x1.7706f5610d8d0p-1, -0x1.3ecc460ef5f50p-1,
  0x1.75f50b522b17cp-1, -0x1.41b941cce0beep-1,    0x1.74e4b040174e5p-1, -0x1.44a41b463c47bp-1,
  0x1.73d5e0c5899f7p-1, -0x1.478cd5959b3d8p-1,    0x1.72c899870f91fp-1, -0x1.4a7373cecf997p-1,
  0x1.71bcd732e940ap-1, -0x1.4d57f8fefe27fp-1,    0x1.70b29680e66fap-1, -0x1.503a682cb1cb3p-1,
  0x1.6fa9d43244380p-1, -0x1.531ac457ee77fp-1,    0x1.6ea28d118b474p-1, -0x1.55f9107a43ee2p-1,
  0x1.6d9cbdf26eaefp-1, -0x1.58d54f86e02f3p-1,    0x1.6c9863b1ab429p-1, -0x1.5baf846aa1b1ap-1,
  0x1.6b957b34e7803p-1, -0x1.5e87b20c2954ap-1,    0x1.6a94016a94017p-1, -0x1.615ddb4bec13cp-1,
  0x1.6993f349cc726p-1, 0x1.61965cdb02c1ep-1,    0x1.68954dd2390bap-1, 0x1.5ec433d5c35aep-1,
  0x1.67980e0bf08c7p-1, 0x1.5bf406b543db1p-1,    0x1.669c31075ab40p-1, 0x1.5925d2b112a59p-1,
  0x1.65a1b3dd13357p-1, 0x1.565995069514cp-1,    0x1.64a893adcd25fp-1, 0x1.538f4af8f72fcp-1,
  0x1.63b0cda236e1cp-1, 0x1.50c6f1d11b97bp-1,    0x1.62ba5eeade65ep-1, 0x1.4e0086dd8baccp-1,
  0x1.61c544c0161c5p-1, 0x1.4b3c077267e9ap-1,    0x1.60d17c61da198p-1, 0x1.487970e958771p-1,
  0x1.5fdf0317b5c6fp-1, 0x1.45b8c0a17df12p-1,    0x1.5eedd630a9fb3p-1, 0x1.42f9f3ff62641p-1,
  0x1.5dfdf303137b6p-1, 0x1.403d086cea79bp-1,    0x1.5d0f56ec91e57p-1, 0x1.3d81fb5946dbcp-1,
  0x1.5c21ff51ef005p-1, 0x1.3ac8ca38e5c5dp-1,    0x1.5b35e99f06714p-1, 0x1.3811728564cb2p-1,
  0x1.5a4b1346add2bp-1, 0x1.355bf1bd82c8bp-1,    0x1.596179c29d2cep-1, 0x1.32a84565120a9p-1,
  0x1.58791a9357ccep-1, 0x1.2ff66b04ea9d5p-1,    0x1.5791f34015792p-1, 0x1.2d46602adccefp-1,
  0x1.56ac0156ac015p-1, 0x1.2a982269a3dbep-1,    0x1.55c7426b79286p-1, 0x1.27ebaf58d8c9cp-1,
  0x1.54e3b4194ce66p-1, 0x1.25410494e56c8p-1,    0x1.5401540154015p-1, 0x1.22981fbef797ap-1,
  0x1.53201fcb02fb1p-1, 0x1.1ff0fe7cf47a9p-1,    0x1.5240152401524p-1, 0x1.1d4b9e796c245p-1,
  0x1.516131c015161p-1, 0x1.1aa7fd638d33ep-1,    0x1.508373590ec9cp-1, 0x1.180618ef18adep-1,
  0x1.4fa6d7aeb597cp-1, 0x1.1565eed455fc2p-1,    0x1.4ecb5c86b3d24p-1, 0x1.12c77cd00713cp-1,
  0x1.4df0ffac83c01p-1, 0x1.102ac0a35cc1bp-1,    0x1.4d17bef15cb4ep-1, 0x1.0d8fb813eb1efp-1,
  0x1.4c3f982c20723p-1, 0x1.0af660eb9e278p-1,    0x1.4b68893948d1cp-1, 0x1.085eb8f8ae799p-1,
  0x1.4a928ffad5b5cp-1, 0x1.05c8be0d9635ap-1,    0x1.49bdaa583b401p-1, 0x1.03346e0106062p-1,
  0x1.48e9d63e504d1p-1, 0x1.00a1c6adda472p-1,    0x1.4817119f3d325p-1, 0x1.fc218be620a5fp-2,
  0x1.47455a726abf2p-1, 0x1.f702d36777df0p-2,    0x1.4674aeb4717e9p-1, 0x1.f1e75fadf9bdep-2,
  0x1.45a50c670938fp-1, 0x1.eccf2c8fe920bp-2,    0x1.44d67190f8b43p-1, 0x1.e7ba35eb77e2ap-2,
  0x1.4408dc3e05b22p-1, 0x1.e2a877a6b2c0fp-2,    0x1.433c4a7ee52b4p-1, 0x1.dd99edaf6d7e9p-2,
  0x1.4270ba692bc4dp-1, 0x1.d88e93fb2f451p-2,    0x1.41a62a173e821p-1, 0x1.d38666871f467p-2,
  0x1.40dc97a843ae8p-1, 0x1.ce816157f1985p-2,    0x1.4014014014014p-1, 0x1.c97f8079d44ecp-2,
  0x1.3f4c65072bf74p-1, 0x1.c480c0005cccfp-2,    0x1.3e85c12a9d651p-1, 0x1.bf851c067555cp-2,
  0x1.3dc013dc013dcp-1, 0x1.ba8c90ae4ad19p-2,    0x1.3cfb5b51698ebp-1, 0x1.b5971a213acd9p-2,
  0x1.3c3795c553afbp-1, 0x1.b0a4b48fc1b44p-2,    0x1.3b74c1769aa5cp-1, 0x1.abb55c31693aep-2,
  0x1.3ab2dca869b81p-1, 0x1.a6c90d44b704cp-2,    0x1.39f1e5a22f36ep-1, 0x1.a1dfc40f1b7f1p-2,
  0x1.3931daaf8f721p-1, 0x1.9cf97cdce0ec1p-2,    0x1.3872ba2057e04p-1, 0x1.981634011aa74p-2,
  0x1.37b4824872744p-1, 0x1.9335e5d594985p-2,    0x1.36f7317fd9212p-1, 0x1.8e588ebac2dc1p-2,
  0x1.363ac622898b1p-1, 0x1.897e2b17b19a6p-2,    0x1.357f3e9078e5bp-1, 0x1.84a6b759f512dp-2,
  0x1.34c4992d87fd9p-1, 0x1.7fd22ff599d4cp-2,    0x1.340ad461776d3p-1, 0x1.7b0091651528bp-2,
  0x1.3351ee97dbfc6p-1, 0x1.7631d82935a84p-2,    0x1.3299e6401329ap-1, 0x1.716600c914055p-2,
  0x1.31e2b9cd37dc2p-1, 0x1.6c9d07d203fc4p-2,    0x1.312c67b6173eep-1, 0x1.67d6e9d785770p-2,
  0x1.3076ee7525c2cp-1, 0x1.6313a37335d76p-2,    0x1.2fc24c8874486p-1, 0x1.5e533144c1718p-2,
  0x1.2f0e8071a5703p-1, 0x1.59958ff1d52f4p-2,    0x1.2e5b88b5e3104p-1, 0x1.54dabc26105d3p-2,
  0x1.2da963ddd3cfbp-1, 0x1.5022b292f6a45p-2,    0x1.2cf8107590e67p-1, 0x1.4b6d6fefe22a5p-2,
  0x1.2c478d0c9c013p-1, 0x1.46baf0f9f5db8p-2,    0x1.2b97d835d548ep-1, 0x1.420b32740fdd6p-2,
  0x1.2ae8f087718d0p-1, 0x1.3d5e3126bc281p-2,    0x1.2a3ad49af0907p-1, 0x1.38b3e9e027477p-2,
  0x1.298d830d13780p-1, 0x1.340c59741142dp-2,    0x1.28e0fa7dd35a3p-1, 0x1.2f677cbbc0a98p-2,
  0x1.2835399057efdp-1, 0x1.2ac55095f5c5bp-2,    0x1.278a3eeaee650p-1, 0x1.2625d1e6ddf55p-2,
  0x1.26e009370049cp-1, 0x1.2188fd9807266p-2,    0x1.263697210aa18p-1, 0x1.1ceed09853755p-2,
  0x1.258de75895121p-1, 0x1.185747dbecf34p-2,    0x1.24e5f89029305p-1, 0x1.13c2605c398bfp-2,
  0x1.243ec97d49eaep-1, 0x1.0f301717cf0fbp-2,    0x1.239858d86b11fp-1, 0x1.0aa06912675d5p-2,
  0x1.22f2a55ce8fc5p-1, 0x1.06135354d4b19p-2,    0x1.224dadc900489p-1, 0x1.0188d2ecf613ep-2,
  0x1.21a970ddc5ba7p-1, 0x1.fa01c9db57ce7p-3,    0x1.2105ed5f1e336p-1, 0x1.f0f70cdd992e4p-3,
  0x1.20632213b6c6dp-1, 0x1.e7f1691a32d3ap-3,    0x1.1fc10dc4fce8bp-1, 0x1.def0d8d466dbbp-3,
  0x1.1f1faf3f16b64p-1, 0x1.d5f55659210e1p-3,    0x1.1e7f0550db594p-1, 0x1.ccfedbfee13a8p-3,
  0x1.1ddf0ecbcb841p-1, 0x1.c40d6425a5cb4p-3,    0x1.1d3fca840a074p-1, 0x1.bb20e936d6976p-3,
  0x1.1ca13750547fep-1, 0x1.b23965a52ff04p-3,    0x1.1c035409fc1dfp-1, 0x1.a956d3ecade60p-3,
  0x1.1b661f8cde833p-1, 0x1.a0792e9277cadp-3,    0x1.1ac998b75eb90p-1, 0x1.97a07024cbe6ep-3,
  0x1.1a2dbe6a5e3e4p-1, 0x1.8ecc933aeb6e2p-3,    0x1.19928f89362b7p-1, 0x1.85fd927506a46p-3,
  0x1.18f80af9b06dcp-1, 0x1.7d33687c293c8p-3,    0x1.185e2fa401186p-1, 0x1.746e100226edbp-3,
  0x1.17c4fc72bfcb9p-1, 0x1.6bad83c1883bap-3,    0x1.172c7052e1316p-1, 0x1.62f1be7d7774ap-3,
  0x1.16948a33b08fap-1, 0x1.5a3abb01ade21p-3,    0x1.15fd4906c96f1p-1, 0x1.5188742261311p-3,
  0x1.1566abc011567p-1, 0x1.48dae4bc3101dp-3,    0x1.14d0b155b19aep-1, 0x1.403207b414b79p-3,
  0x1.143b58c01143bp-1, 0x1.378dd7f74970fp-3,    0x1.13a6a0f9cf01ep-1, 0x1.2eee507b402ffp-3,
  0x1.131288ffbb3b6p-1, 0x1.26536c3d8c36cp-3,    0x1.127f0fd0d2295p-1, 0x1.1dbd2643d1913p-3,
  0x1.11ec346e36092p-1, 0x1.152b799bb3cd0p-3,    0x1.1159f5db29606p-1, 0x1.0c9e615ac4e19p-3,
  0x1.10c8531d0952ep-1, 0x1.0415d89e7444bp-3,    0x1.10374b3b480aap-1, 0x1.f723b517fc51fp-4,
  0x1.0fa6dd3f67322p-1, 0x1.e624c4a0b5e15p-4,    0x1.0f170834f27fap-1, 0x1.d52ed6405d87ap-4,
  0x1.0e87cb297a51ep-1, 0x1.c441e06f72a93p-4,    0x1.0df9252c8e5e6p-1, 0x1.b35dd9b58baa8p-4,
  0x1.0d6b154fb86f9p-1, 0x1.a282b8a936174p-4,    0x1.0cdd9aa677344p-1, 0x1.91b073efd7314p-4,
  0x1.0c50b446391f3p-1, 0x1.80e7023d8ccc8p-4,    0x1.0bc4614657569p-1, 0x1.70265a550e77bp-4,
  0x1.0b38a0c010b39p-1, 0x1.5f6e73078efc3p-4,    0x1.0aad71ce84d16p-1, 0x1.4ebf43349e26ap-4,
  0x1.0a22d38eaf2bfp-1, 0x1.3e18c1ca0ae99p-4,    0x1.0998c51f624d5p-1, 0x1.2d7ae5c3c5bb7p-4,
  0x1.090f45a1430aap-1, 0x1.1ce5a62bc3540p-4,    0x1.08865436c3cf7p-1, 0x1.0c58fa19dfaabp-4,
  0x1.07fdf0041ff7cp-1, 0x1.f7a9b16782855p-5,    0x1.0776182f57386p-1, 0x1.d6b272597981fp-5,
  0x1.06eecbe029155p-1, 0x1.b5cc258b718e7p-5,    0x1.06680a4010668p-1, 0x1.94f6b99a24473p-5,
  0x1.05e1d27a3ee9cp-1, 0x1.74321d3d006d2p-5,    0x1.055c23bb98e2ap-1, 0x1.537e3f45f354ep-5,
  0x1.04d6fd32b0c7bp-1, 0x1.32db0ea132e10p-5,    0x1.04525e0fc2fcbp-1, 0x1.12487a5507f68p-5,
  0x1.03ce4584b19a0p-1, 0x1.e38ce30333100p-6,    0x1.034ab2c50040dp-1, 0x1.a2a9c6c17044dp-6,
  0x1.02c7a505cffbfp-1, 0x1.61e77e8b53f9fp-6,    0x1.02451b7ddb2d2p-1, 0x1.2145e939ef1bcp-6,
  0x1.01c315657186bp-1, 0x1.c189cbb0e283fp-7,    0x1.014191f674111p-1, 0x1.40c8a7478788dp-7,
  0x1.00c0906c513cfp-1, 0x1.809048289860ap-8,    0x1.0000000000000p-1, 0.0,
};
// -2 * log1p(r) = r * (c1 + r * (c2 + ... + r * c7)),  c_k = -2 * (-1)^(k+1) / k
RM_MC_CONST double kNeg2Log1p[7] = {-2.0, 1.0, -2.0 / 3, 0.5, -0.4, 1.0 / 3, -2.0 / 7};
// fdlibm k_sin.c S1..S6, k_cos.c C1..C6
RM_MC_CONST double kSinK[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                               -2.50507602534068634195e-08, 1.58969099521155010221e-10};
RM_MC_CONST double kCosK[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
                               2.08757232129817482790e-09, -1.13596475577881948265e-11};

RM_MC_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

// -2 * log(max(x * 2^-53, f64::MIN_POSITIVE)) for x < 2^53 (the squared Box-Muller radius); tab = kNeg2LogTab or a copy of it.
RM_MC_HD double neg2log_u53(uint64_t x, const double* tab) {
  if (x == 0) return 1416.7928370645282;  // -2 * log(2^-1022): the host clamps u1 = 0 to f64::MIN_POSITIVE (random.rs:279-281)
  const int lz = clz64(x);                 // 11 .. 63
  const uint64_t mant = x << (lz - 11);    // bit 52 set: m = mant * 2^-52 in [1, 2), u = m * 2^(10 - lz)
  const int i = (int)(mant >> 44) & 0xFF;
  const double m = bits_to_double((mant & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
  const double ed = (double)(10 - lz + (i >= 106 ? 1 : 0));
  const double r = fma(m, tab[2 * i], -1.0);  // one rounding of an exact cancellation: relative error 2^-53 of r
  double p = kNeg2Log1p[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 5; k >= 0; --k) p = fma(p, r, kNeg2Log1p[k]);
  double w = fma(ed, -2.0 * 6.93147180369123816490e-01, tab[2 * i + 1]);  // e * (2 ln2_hi) is exact (32-bit constant)
  w = fma(p, r, w);
  return fma(ed, -2.0 * 1.90821492927058770002e-10, w);
}

// sin and cos of 2*pi*(x * 2^-53): quarter turns and the reduced angle by integer arithmetic, minimax kernels on |a| <= pi/4.
RM_MC_HD void sincos_turn_u53(uint64_t x, double* sn, double* cs) {
  const unsigned q = (unsigned)((x + (1ull << 50)) >> 51);            // nearest quarter turn, 0 .. 4
  const int64_t f = (int64_t)(x - ((uint64_t)q << 51));               // [-2^50, 2^50): exact in a double
  const double a = (double)f * 6.975736996017264e-16;                 // * (pi/2) * 2^-51 (the power of two is exact)
  const double z = a * a;
  double ps = kSinK[5], pc = kCosK[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 4; k >= 1; --k) ps = fma(ps, z, kSinK[k]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 4; k >= 0; --k) pc = fma(pc, z, kCosK[k]);
  const double sx = fma(z * a, fma(z, ps, kSinK[0]), a);              // a + a^3 (S1 + z (S2 + ...))
  const double cx = fma(z, fma(z, pc, -0.5), 1.0);                    // 1 - z/2 + z^2 (C1 + ...)
  const bool swap = (q & 1u) != 0;
  const uint64_t s0 = double_to_bits(swap ? cx : sx), c0 = double_to_bits(swap ? sx : cx);
  // quadrants: 0 (s, c)  1 (c, -s)  2 (-s, -c)  3 (-c, s)  4 = 0: the sign flips are XORs on the sign bit
  *sn = bits_to_double(s0 ^ ((uint64_t)(q & 2u) << 62));
  *cs = bits_to_double(c0 ^ ((uint64_t)((q + 1u) & 2u) << 62));
}

}  // namespace rm_mc
