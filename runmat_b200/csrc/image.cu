// image.cu — image_normalize (row a12), imfilter (row a13) and the value+index dim reductions (row a6).
//
// image_normalize reference semantics: crates/runmat-accelerate/src/simple_provider.rs:7893-7994
//   per image b of a [B,H,W] column-major tensor (batch is the STRIDE-1 axis: stride_h = B, stride_w = B*H):
//   mu = sum(x)/P; sigma = sqrt(sum((x-mu)^2)/P + eps); y = (x-mu)*(1/sigma) [*gain] [+bias] [max 0] [pow gamma].
// The wgpu provider streams this in batch windows with three sweeps (shaders/image_normalize.rs,
// provider/helpers.rs:96-380). B200 design: because batch is stride-1 the whole tensor is one contiguous
// [B x P] matrix, so both sweeps are fully coalesced flat streams:
//   sweep 1  per-image shifted moments (sum(x-K), sum((x-K)^2), K = first pixel) accumulated in f64 —
//            algebraically the two-pass population variance, without a second statistics pass;
//            per-block partials are combined in fixed order by a tiny finalize kernel (deterministic);
//   sweep 2  normalise + gain/bias/clamp/gamma, one read + one write.
// HBM traffic: 2 reads + 1 write per pixel (12 B/px f32); a 4K batch is far larger than L2, so the second
// read cannot be served on-chip with a batch-fastest layout (DESIGN.md "image_normalize").
#include <cuda.h>

#include <type_traits>

#include "common.h"

namespace rm {

namespace {

// ---- sweep 1: thread (b, lane) = (tid % B, tid / B); each lane strides over pixels ------------------------------
// Works for any B <= 1024: consecutive threads read consecutive addresses (b fastest), fixed image per thread.
template <typename T>
__global__ void __launch_bounds__(1024)
moments_partial_kernel(const T* __restrict__ x, uint64_t B, uint64_t P, double* __restrict__ partial /*[grid][2][B]*/) {
  extern __shared__ double sh[];  // [lanes][2][B]
  const uint32_t lanes = blockDim.x / (uint32_t)B;  // >= 1
  const uint32_t b = threadIdx.x % (uint32_t)B, lane = threadIdx.x / (uint32_t)B;
  double s1 = 0.0, s2 = 0.0;
  if (lane < lanes) {
    const double K = (double)x[b];  // pixel 0 of image b
    const uint64_t stride = (uint64_t)gridDim.x * lanes;
    uint64_t p = (uint64_t)blockIdx.x * lanes + lane;
    // 4 independent loads in flight per thread
    for (; p + 3 * stride < P; p += 4 * stride) {
      const double d0 = (double)x[b + p * B] - K, d1 = (double)x[b + (p + stride) * B] - K;
      const double d2 = (double)x[b + (p + 2 * stride) * B] - K, d3 = (double)x[b + (p + 3 * stride) * B] - K;
      s1 += (d0 + d1) + (d2 + d3);
      s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    for (; p < P; p += stride) { const double d = (double)x[b + p * B] - K; s1 += d; s2 += d * d; }
    sh[(lane * 2 + 0) * B + b] = s1;
    sh[(lane * 2 + 1) * B + b] = s2;
  }
  __syncthreads();
  if (threadIdx.x < B) {
    double a1 = 0.0, a2 = 0.0;
    for (uint32_t l = 0; l < lanes; ++l) { a1 += sh[(l * 2 + 0) * B + threadIdx.x]; a2 += sh[(l * 2 + 1) * B + threadIdx.x]; }
    partial[((uint64_t)blockIdx.x * 2 + 0) * B + threadIdx.x] = a1;
    partial[((uint64_t)blockIdx.x * 2 + 1) * B + threadIdx.x] = a2;
  }
}

// Fast path (B %% VEC == 0 and the grid stride keeps each lane on a fixed image): 16-byte vector loads, 4 in flight per
// thread, per-lane f64 accumulators, one shared-memory atomic per lane at the end.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
moments_partial_vec_kernel(const T* __restrict__ x, uint32_t B, uint64_t total, double* __restrict__ partial /*[grid][2][B]*/) {
  extern __shared__ double sh[];  // [2][B]
  for (uint32_t i = threadIdx.x; i < 2 * B; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const uint64_t nvec = total / VEC, nthr = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b0 = (uint32_t)((v0 * VEC) % B);  // fixed for this thread: (nthr*VEC) %% B == 0
  double K[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { K[l] = (double)x[b0 + l]; s1[l] = 0.0; s2[l] = 0.0; }
  typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type V;
  const V* xv = reinterpret_cast<const V*>(x);
  uint64_t v = v0;
  for (; v + 7 * nthr < nvec; v += 8 * nthr) {  // 8 x 16 B in flight per thread (r02 ncu: 4 were latency-bound at 2.85 TB/s)
    V a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = __ldcs(xv + v + (uint64_t)u * nthr);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const T* e = reinterpret_cast<const T*>(&a[u]);
#pragma unroll
      for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
    }
  }
  for (; v < nvec; v += nthr) {
    const V a = __ldcs(xv + v);
    const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
    for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
  }
#pragma unroll
  for (int l = 0; l < VEC; ++l) { atomicAdd(&sh[b0 + l], s1[l]); atomicAdd(&sh[B + b0 + l], s2[l]); }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < B; i += blockDim.x) {
    partial[((uint64_t)blockIdx.x * 2 + 0) * B + i] = sh[i];
    partial[((uint64_t)blockIdx.x * 2 + 1) * B + i] = sh[B + i];
  }
}

// Same sweep with a deterministic, atomic-free finish (r04 harness, scripts/img_dev/moments_variants.cu: 41.4 us vs 88.3 us on
// [8,2160,3840] f32; the atomic version ended every thread with 2*VEC shared-memory f64 atomics onto 2*B addresses, 128-way
// contended, and summed in arrival order). Consecutive threads own consecutive vectors, so the lanes of a warp cycle through
// the G = B / VEC image groups with period G. For G a power of two <= 32 an xor-butterfly over the offsets 16 .. G folds exactly
// the lanes that share a group; lanes 0 .. G-1 then publish per-warp slots and the slots are folded in warp order.
// KEEP: plain read-only loads instead of evict-first streaming loads, so the end of the tensor is still in L2 when the normalise
// sweep starts there (it runs end-to-front).
template <typename T, int VEC, bool KEEP>
__global__ void __launch_bounds__(256)
moments_partial_fold_kernel(const T* __restrict__ x, uint32_t B, uint64_t total, double* __restrict__ partial /*[grid][2][B]*/) {
  extern __shared__ double shw[];  // [8 warps][2][B]
  pdl_prologue();
  const uint64_t nvec = total / VEC, nthr = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b0 = (uint32_t)((v0 * VEC) % B);  // fixed for this thread: (nthr*VEC) %% B == 0
  double K[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { K[l] = (double)x[b0 + l]; s1[l] = 0.0; s2[l] = 0.0; }
  typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type V;
  const V* xv = reinterpret_cast<const V*>(x);
  uint64_t v = v0;
  for (; v + 7 * nthr < nvec; v += 8 * nthr) {
    V a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = KEEP ? __ldg(xv + v + (uint64_t)u * nthr) : __ldcs(xv + v + (uint64_t)u * nthr);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const T* e = reinterpret_cast<const T*>(&a[u]);
#pragma unroll
      for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
    }
  }
  for (; v < nvec; v += nthr) {
    const V a = KEEP ? __ldg(xv + v) : __ldcs(xv + v);
    const T* e = reinterpret_cast<const T*>(&a);
#pragma unroll
    for (int l = 0; l < VEC; ++l) { const double d = (double)e[l] - K[l]; s1[l] += d; s2[l] += d * d; }
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = B / VEC;
#pragma unroll
  for (int l = 0; l < VEC; ++l)
    for (uint32_t off = 16; off >= G && off > 0; off >>= 1) {
      s1[l] += __shfl_xor_sync(0xffffffffu, s1[l], off);
      s2[l] += __shfl_xor_sync(0xffffffffu, s2[l], off);
    }
  if (lane < G) {
#pragma unroll
    for (int l = 0; l < VEC; ++l) { shw[(warp * 2 + 0) * B + b0 + l] = s1[l]; shw[(warp * 2 + 1) * B + b0 + l] = s2[l]; }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 2 * B; i += blockDim.x) {
    const uint32_t m = i / B, b = i % B;
    double t = 0.0;
    for (uint32_t w = 0; w < 8; ++w) t += shw[(w * 2 + m) * B + b];  // warp order: deterministic
    partial[((uint64_t)blockIdx.x * 2 + m) * B + b] = t;
  }
}

// stats[b] = {mean, inv_sigma}. One CTA per image: threads stride over the per-block partials, fixed-order tree in
// shared memory (deterministic).
template <typename T>
__global__ void __launch_bounds__(128)
moments_finalize_kernel(const T* __restrict__ x, const double* __restrict__ partial, uint32_t nblocks, uint64_t B, uint64_t P,
                        double eps, double* __restrict__ stats) {
  __shared__ double r1[128], r2[128];
  pdl_prologue();
  const uint64_t b = blockIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t i = threadIdx.x; i < nblocks; i += 128) { s1 += partial[((uint64_t)i * 2 + 0) * B + b]; s2 += partial[((uint64_t)i * 2 + 1) * B + b]; }
  r1[threadIdx.x] = s1;
  r2[threadIdx.x] = s2;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) { r1[threadIdx.x] += r1[threadIdx.x + off]; r2[threadIdx.x] += r2[threadIdx.x + off]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    s1 = r1[0];
    s2 = r2[0];
    const double K = (double)x[b];
    const double n = (double)P;
    const double dm = s1 / n;                       // mean of (x-K)
    double var = (s2 - s1 * dm) / n;                // sum((x-mean)^2)/P
    if (var < 0.0) var = 0.0;
    const double sigma = sqrt(var + eps);
    stats[2 * b] = K + dm;
    stats[2 * b + 1] = sigma > 0.0 ? 1.0 / sigma : 0.0;  // simple_provider.rs:7962
  }
}

struct NormParams {
  int has_gain, has_bias, has_gamma, clamp_zero;
  double gain, bias, gamma;
};

// f32 storage: pow as WGSL defines it, exp2(y * log2(x)) with the full-precision (non-approximate) log2f/exp2f — the
// accuracy class of the reference's own f32 shader (shaders/image_normalize.rs) at a fraction of powf's cost
// (powf carries double-float internals; ncu r02: the normalise sweep was compute-bound on it, SM throughput 81 %).
// The shortcut is only taken for a >= 0 (what clamp_zero guarantees); a negative base (gamma without the clamp) goes
// through powf so that e.g. (-x)^2 is x^2 and not NaN, like the host's f32::powf and the generic kernel below.
__device__ __forceinline__ float rm_powT(float a, float b) {
  if (b == 0.0f) return 1.0f;
  if (a < 0.0f) return powf(a, b);
  return exp2f(b * log2f(a));
}
__device__ __forceinline__ double rm_powT(double a, double b) { return pow(a, b); }

// Fixed-lane variant: (gridDim*blockDim*VEC) %% B == 0, so each thread's lane l always belongs to image (b0+l): the
// per-image mean / 1/sigma live in registers and the inner loop is load -> 4 FLOPs (+pow) -> store.
// f32 clamp + gamma specialisation. ncu r06 ([64,2160,3840] f32): the generic body below issues ~50 instructions per pixel
// (per-element option tests, the a < 0 guard, the denormal pre/post-scaling inside log2f/exp2f) and runs at 4.5 TB/s with the
// issue slots 79 % busy. With the options fixed at compile time and the clamp guaranteeing a >= 0 the pixel is
// sub, mul, mul, (add), max, MUFU.LG2, mul, MUFU.EX2 and the sweep is an HBM stream again. `.ftz` on the two MUFU ops: a
// clamped value below 2^-126 is treated as 0 (what the WGSL pow of the wgpu provider is allowed to do as well); 0^g = 0,
// inf^g = inf, 1^g = 1 come out of lg2/ex2 directly; gamma == 0 never reaches this kernel.
__device__ __forceinline__ float pow_nonneg_f32(float a, float g) {
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(a));
  l *= g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
  return r;
}
// VEC = 8 (B % 8 == 0): 256-bit accesses like the fused elementwise kernel (LDG.E.256 / STG.E.256, L1 no-allocate); VEC = 4 otherwise.
struct __align__(32) f32x8 { float x[8]; };
template <int VEC>
__device__ __forceinline__ void ld_stream_f32(const float* p, float* e) {
  if (VEC == 8)
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(e[0]), "=f"(e[1]), "=f"(e[2]), "=f"(e[3]), "=f"(e[4]), "=f"(e[5]), "=f"(e[6]), "=f"(e[7]) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e[0]), "=f"(e[1]), "=f"(e[2]), "=f"(e[3]) : "l"(p));
}
template <int VEC>
__device__ __forceinline__ void st_stream_f32(float* p, const float* e) {
  if (VEC == 8)
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]), "f"(e[4]), "f"(e[5]), "f"(e[6]), "f"(e[7]) : "memory");
  else
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3]) : "memory");
}
template <int VEC>
__global__ void __launch_bounds__(256)
normalize_fixed_clampgamma_f32_kernel(const float* __restrict__ x, float* __restrict__ y, uint32_t B, uint64_t total, const double* __restrict__ stats,
                                      const __grid_constant__ NormParams np, int rev) {
  constexpr int U = VEC == 8 ? 2 : 4;  // 64 bytes per thread in flight
  pdl_prologue();
  const uint64_t nvec = total / VEC, nthr = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // rev: sweep from the end of the tensor to the front (vector v -> nvec-1-v). The moments sweep has just read the tensor front to
  // end, so its last ~100 MB are still in the 126 MB L2 and serve the first reads here. nthr*VEC %% B == 0 keeps the image fixed.
  auto at = [&](uint64_t v) { return rev ? nvec - 1 - v : v; };
  const uint32_t b0 = (uint32_t)((at(v0 < nvec ? v0 : 0) * VEC) % B);
  float mean[VEC], inv[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { mean[l] = (float)stats[2 * (b0 + l)]; inv[l] = (float)stats[2 * (b0 + l) + 1]; }
  const float gain = np.has_gain ? (float)np.gain : 1.0f, bias = (float)np.bias, gamma = (float)np.gamma;  // v * 1.0f is the identity, bit for bit
  const bool has_bias = np.has_bias != 0;
  auto body = [&](float* e) {
#pragma unroll
    for (int l = 0; l < VEC; ++l) {
      float val = (e[l] - mean[l]) * inv[l];
      val *= gain;
      if (has_bias) val += bias;
      val = val > 0.0f ? val : 0.0f;  // f64::max(v, 0.0): NaN -> 0
      e[l] = pow_nonneg_f32(val, gamma);
    }
  };
  uint64_t v = v0;
  for (; v + (U - 1) * nthr < nvec; v += U * nthr) {
    float a[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) ld_stream_f32<VEC>(x + at(v + (uint64_t)u * nthr) * VEC, a[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) { body(a[u]); st_stream_f32<VEC>(y + at(v + (uint64_t)u * nthr) * VEC, a[u]); }
  }
  for (; v < nvec; v += nthr) {
    float a[VEC];
    ld_stream_f32<VEC>(x + at(v) * VEC, a);
    body(a);
    st_stream_f32<VEC>(y + at(v) * VEC, a);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_fixed_kernel(const T* __restrict__ x, T* __restrict__ y, uint32_t B, uint64_t total, const double* __restrict__ stats,
                       const __grid_constant__ NormParams np, int rev) {
  typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type V;
  pdl_prologue();
  const uint64_t nvec = total / VEC, nthr = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t v0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto at = [&](uint64_t v) { return rev ? nvec - 1 - v : v; };  // end-to-front sweep: see normalize_fixed_clampgamma_f32_kernel
  const uint32_t b0 = (uint32_t)((at(v0 < nvec ? v0 : 0) * VEC) % B);
  T mean[VEC], inv[VEC];
#pragma unroll
  for (int l = 0; l < VEC; ++l) { mean[l] = (T)stats[2 * (b0 + l)]; inv[l] = (T)stats[2 * (b0 + l) + 1]; }
  const T gain = (T)np.gain, bias = (T)np.bias, gamma = (T)np.gamma;
  const V* xv = reinterpret_cast<const V*>(x);
  V* yv = reinterpret_cast<V*>(y);
  auto body = [&](V a) {
    T* e = reinterpret_cast<T*>(&a);
#pragma unroll
    for (int l = 0; l < VEC; ++l) {
      T val = (e[l] - mean[l]) * inv[l];
      if (np.has_gain) val *= gain;
      if (np.has_bias) val += bias;
      if (np.clamp_zero) val = val > (T)0 ? val : (T)0;
      if (np.has_gamma) val = rm_powT(val, gamma);
      e[l] = val;
    }
    return a;
  };
  uint64_t v = v0;
  for (; v + nthr < nvec; v += 2 * nthr) {
    const V a = __ldcs(xv + at(v)), b = __ldcs(xv + at(v + nthr));
    __stcs(yv + at(v), body(a));
    __stcs(yv + at(v + nthr), body(b));
  }
  for (; v < nvec; v += nthr) __stcs(yv + at(v), body(__ldcs(xv + at(v))));
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
normalize_kernel(const T* __restrict__ x, T* __restrict__ y, uint64_t B, uint64_t total, const double* __restrict__ stats,
                 const __grid_constant__ NormParams np) {
  // flat stream; image index = e % B. VEC elements per thread per iteration (16-byte accesses when VEC>1).
  const uint64_t nvec = total / VEC;
  const uint64_t nthr = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += nthr) {
    T v[VEC];
    if (VEC == 4) *reinterpret_cast<float4*>(v) = __ldcs(reinterpret_cast<const float4*>(x) + i);
    else if (VEC == 2) *reinterpret_cast<double2*>(v) = __ldcs(reinterpret_cast<const double2*>(x) + i);
    else v[0] = x[i];
    const uint64_t e0 = i * VEC;
    uint32_t b = (uint32_t)(e0 % B);
#pragma unroll
    for (int l = 0; l < VEC; ++l) {
      const T mean = (T)stats[2 * b], inv = (T)stats[2 * b + 1];
      T val = (v[l] - mean) * inv;
      if (np.has_gain) val *= (T)np.gain;
      if (np.has_bias) val += (T)np.bias;
      if (np.clamp_zero) val = val > (T)0 ? val : (T)0;  // f64::max(v, 0.0): NaN -> 0
      if (np.has_gamma) val = rm_powT(val, (T)np.gamma);  // the same pow as the fixed-lane kernel: the result must not depend on the path taken
      v[l] = val;
      if (++b == B) b = 0;
    }
    if (VEC == 4) __stcs(reinterpret_cast<float4*>(y) + i, *reinterpret_cast<float4*>(v));
    else if (VEC == 2) __stcs(reinterpret_cast<double2*>(y) + i, *reinterpret_cast<double2*>(v));
    else y[i] = v[0];
  }
  // ragged tail
  for (uint64_t e = nvec * VEC + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += nthr) {
    const uint32_t b = (uint32_t)(e % B);
    const T mean = (T)stats[2 * b], inv = (T)stats[2 * b + 1];
    T val = (x[e] - mean) * inv;
    if (np.has_gain) val *= (T)np.gain;
    if (np.has_bias) val += (T)np.bias;
    if (np.clamp_zero) val = val > (T)0 ? val : (T)0;
    if (np.has_gamma) val = sizeof(T) == 4 ? (T)powf((float)val, (float)np.gamma) : (T)pow((double)val, np.gamma);
    y[e] = val;
  }
}

// ---- imfilter (image/filters/imfilter.rs:476-745) ---------------------------------------------------------------------
struct FilterParams {
  uint64_t ie[3], ke[3], oe[3];
  int64_t origin[3], base[3];
  int padding, mode;
  double cval;
  // 1.0f, but only known at run time: the packed f32 kernel accumulates with fma.rn.f32x2(acc, one, product), which is
  // bit-identical to add.rn (x*1 is exact) and which ptxas cannot contract with the preceding mul.rn.f32x2 into an FFMA2 the
  // way it contracts mul.rn.f32x2 + add.rn.f32x2 (even under -fmad=false) or a literal 1.0f (it folds that back to an add).
  float one = 1.0f;
};
__device__ __forceinline__ int64_t clamp_index(int64_t c, int64_t len) { return (len <= 0 || c <= 0) ? 0 : (c >= len ? len - 1 : c); }
__device__ __forceinline__ int64_t wrap_index(int64_t c, int64_t len) { if (len <= 0) return 0; c %= len; if (c < 0) c += len; return c; }
__device__ __forceinline__ int64_t reflect_index(int64_t c, int64_t len) {
  if (len <= 1) return 0;
  const int64_t period = 2 * len - 2;
  int64_t v = c % period;
  if (v < 0) v += period;
  if (v >= len) v = period - v;
  return v;
}
template <typename T>
__global__ void __launch_bounds__(256)
imfilter_kernel(const T* __restrict__ img, const T* __restrict__ ker, T* __restrict__ out, uint64_t total, const __grid_constant__ FilterParams fp) {
  const uint64_t istr1 = fp.ie[0], istr2 = fp.ie[0] * fp.ie[1];
  const uint64_t kstr1 = fp.ke[0], kstr2 = fp.ke[0] * fp.ke[1];
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (uint64_t)gridDim.x * blockDim.x) {
    const int64_t o0 = (int64_t)(o % fp.oe[0]), o1 = (int64_t)((o / fp.oe[0]) % fp.oe[1]), o2 = (int64_t)(o / (fp.oe[0] * fp.oe[1]));
    T sum = (T)0;
    // kernel points in column-major order; sum += k * sample (imfilter.rs:731-745), no FMA
    for (uint64_t k2 = 0; k2 < fp.ke[2]; ++k2)
      for (uint64_t k1 = 0; k1 < fp.ke[1]; ++k1)
        for (uint64_t k0 = 0; k0 < fp.ke[0]; ++k0) {
          const uint64_t lin = fp.mode == 0 ? k0 + k1 * kstr1 + k2 * kstr2
                                            : (fp.ke[0] - 1 - k0) + (fp.ke[1] - 1 - k1) * kstr1 + (fp.ke[2] - 1 - k2) * kstr2;
          const T kv = __ldg(ker + lin);
          int64_t c[3] = {o0 + fp.base[0] + ((int64_t)k0 - fp.origin[0]), o1 + fp.base[1] + ((int64_t)k1 - fp.origin[1]),
                          o2 + fp.base[2] + ((int64_t)k2 - fp.origin[2])};
          bool constant = false;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int64_t len = (int64_t)fp.ie[d];
            if (c[d] < 0 || c[d] >= len) {
              if (fp.padding == 0) constant = true;
              else c[d] = fp.padding == 1 ? clamp_index(c[d], len) : (fp.padding == 3 ? wrap_index(c[d], len) : reflect_index(c[d], len));
            }
          }
          const T sample = constant ? (T)fp.cval : img[(uint64_t)c[0] + (uint64_t)c[1] * istr1 + (uint64_t)c[2] * istr2];
          sum += kv * sample;
        }
    out[o] = sum;
  }
}

// Tiled variant for 2-D kernels (the common case: fspecial-style filters on H x W x C images). One CTA produces a
// 64 x 16 output tile of one plane: the input tile + halo is staged ONCE in shared memory with the padding rule resolved
// per row/column (separable index maps), then each thread accumulates its 4 outputs from shared memory, visiting the
// kernel points in the host's column-major order with separate multiply and add (-fmad=false): bit-identical to
// imfilter.rs:731-745. HBM traffic ~ (1 + halo) read + 1 write per sample instead of taps x (index math + load).
constexpr int IMF_TX = 64, IMF_TY = 16, IMF_MAXK = 15;
template <typename T>
__global__ void __launch_bounds__(256)
imfilter_tiled_kernel(const T* __restrict__ img, const T* __restrict__ ker, T* __restrict__ out, const __grid_constant__ FilterParams fp) {
  extern __shared__ __align__(16) unsigned char imf_smem[];
  const int K0 = (int)fp.ke[0], K1 = (int)fp.ke[1];
  const int SX = IMF_TX + K0 - 1, SY = IMF_TY + K1 - 1;
  T* tile = reinterpret_cast<T*>(imf_smem);                   // [SY][SX]
  T* w = tile + SX * SY;                                      // [K1][K0] in application order
  int* map0 = reinterpret_cast<int*>(w + K0 * K1);            // [SX] source row (dim 0) or -1
  int* map1 = map0 + SX;                                      // [SY] source col (dim 1) or -1
  const int64_t t0 = (int64_t)blockIdx.x * IMF_TX, t1 = (int64_t)blockIdx.y * IMF_TY;
  const uint64_t plane = blockIdx.z;
  const int tid = threadIdx.x;
  auto remap = [&](int64_t c, int64_t len) -> int {
    if (c >= 0 && c < len) return (int)c;
    if (fp.padding == 0) return -1;
    return (int)(fp.padding == 1 ? clamp_index(c, len) : (fp.padding == 3 ? wrap_index(c, len) : reflect_index(c, len)));
  };
  for (int i = tid; i < SX; i += 256) map0[i] = remap(t0 + i + fp.base[0] - fp.origin[0], (int64_t)fp.ie[0]);
  for (int i = tid; i < SY; i += 256) map1[i] = remap(t1 + i + fp.base[1] - fp.origin[1], (int64_t)fp.ie[1]);
  for (int i = tid; i < K0 * K1; i += 256) {
    const int k0 = i % K0, k1 = i / K0;
    w[i] = fp.mode == 0 ? ker[k0 + k1 * K0] : ker[(K0 - 1 - k0) + (K1 - 1 - k1) * K0];
  }
  __syncthreads();
  const T* src = img + plane * fp.ie[0] * fp.ie[1];
  for (int i = tid; i < SX * SY; i += 256) {
    const int sx = i % SX, sy = i / SX;
    const int r = map0[sx], c = map1[sy];
    tile[i] = (r < 0 || c < 0) ? (T)fp.cval : src[(uint64_t)r + (uint64_t)c * fp.ie[0]];
  }
  __syncthreads();
  const int lx = tid & 63, ly = tid >> 6;  // 4 thread rows; each thread produces outputs ly, ly+4, ly+8, ly+12
  T acc[4] = {(T)0, (T)0, (T)0, (T)0};
  for (int k1 = 0; k1 < K1; ++k1)
    for (int k0 = 0; k0 < K0; ++k0) {
      const T kv = w[k0 + k1 * K0];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += kv * tile[(ly + 4 * j + k1) * SX + lx + k0];
    }
  const uint64_t o0 = (uint64_t)t0 + lx;
  if (o0 < fp.oe[0]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t o1 = (uint64_t)t1 + ly + 4 * j;
      if (o1 < fp.oe[1]) out[o0 + o1 * fp.oe[0] + plane * fp.oe[0] * fp.oe[1]] = acc[j];
    }
  }
}

// Register-blocked specialisation for the common square kernels (3x3, 5x5, 7x7). The generic tiled kernel issues one shared
// load per (tap, output) and is LDS-bound (r03: 0.9 TB/s at 8 B/sample). Here the K0*K1 weights live in registers and each
// thread produces RB = 8 outputs along dim 1: walking the staged input columns j (outer) and k0 (inner), every staged value
// is loaded once and feeds up to K1 outputs. For a fixed output the taps still arrive as k1 ascending, k0 ascending —
// exactly the host's column-major kernel order (imfilter.rs:731-745) — with separate multiply and add, so results stay
// bit-identical to the reference.
constexpr int RB = 8, FX = 64, FY = 32;  // CTA tile: 64 (dim 0) x 32 (dim 1) outputs; 256 threads = 64 x 4 strips of RB
template <typename T, int K0, int K1>
__global__ void __launch_bounds__(256)
imfilter_regblock_kernel(const T* __restrict__ img, const T* __restrict__ ker, T* __restrict__ out, const __grid_constant__ FilterParams fp) {
  constexpr int SX = FX + K0 - 1, SY = FY + K1 - 1;
  __shared__ T tile[SY][SX];
  __shared__ int map0[SX], map1[SY];
  const int64_t t0 = (int64_t)blockIdx.x * FX, t1 = (int64_t)blockIdx.y * FY;
  const uint64_t plane = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto remap = [&](int64_t c, int64_t len) -> int {
    if (c >= 0 && c < len) return (int)c;
    if (fp.padding == 0) return -1;
    return (int)(fp.padding == 1 ? clamp_index(c, len) : (fp.padding == 3 ? wrap_index(c, len) : reflect_index(c, len)));
  };
  // tiles whose halo lies inside the image (all but the border ring) skip the index maps
  const int64_t h0 = t0 + fp.base[0] - fp.origin[0], h1 = t1 + fp.base[1] - fp.origin[1];
  const bool interior = h0 >= 0 && h1 >= 0 && h0 + SX <= (int64_t)fp.ie[0] && h1 + SY <= (int64_t)fp.ie[1];
  if (!interior) {
    for (int i = tid; i < SX; i += 256) map0[i] = remap(h0 + i, (int64_t)fp.ie[0]);
    for (int i = tid; i < SY; i += 256) map1[i] = remap(h1 + i, (int64_t)fp.ie[1]);
  }
  T w[K0 * K1];  // application order (already flipped for convolution)
#pragma unroll
  for (int i = 0; i < K0 * K1; ++i) {
    const int k0 = i % K0, k1 = i / K0;
    w[i] = __ldg(ker + (fp.mode == 0 ? k0 + k1 * K0 : (K0 - 1 - k0) + (K1 - 1 - k1) * K0));
  }
  const T* src = img + plane * fp.ie[0] * fp.ie[1];
  if (interior) {
    const T* col = src + (uint64_t)h0 + (uint64_t)(h1 + warp) * fp.ie[0];
    const uint64_t step = 8 * fp.ie[0];
    for (int sy = warp; sy < SY; sy += 8, col += step) {  // one warp per staged column: coalesced along dim 0
#pragma unroll
      for (int sx = lane; sx < SX; sx += 32) tile[sy][sx] = col[sx];
    }
  } else {
    __syncthreads();
    for (int sy = warp; sy < SY; sy += 8) {
      const int c = map1[sy];
      for (int sx = lane; sx < SX; sx += 32) {
        const int r = map0[sx];
        tile[sy][sx] = (r < 0 || c < 0) ? (T)fp.cval : src[(uint64_t)r + (uint64_t)c * fp.ie[0]];
      }
    }
  }
  __syncthreads();
  const int lx = tid & 63, ly0 = (tid >> 6) * RB;
  T acc[RB];
#pragma unroll
  for (int o = 0; o < RB; ++o) acc[o] = (T)0;
#pragma unroll
  for (int j = 0; j < RB + K1 - 1; ++j) {
#pragma unroll
    for (int k0 = 0; k0 < K0; ++k0) {
      const T v = tile[ly0 + j][lx + k0];
#pragma unroll
      for (int k1 = 0; k1 < K1; ++k1) {
        const int o = j - k1;
        if (o >= 0 && o < RB) acc[o] += w[k0 + k1 * K0] * v;
      }
    }
  }
  const uint64_t o0 = (uint64_t)t0 + lx, o1 = (uint64_t)t1 + ly0;
  if (o0 < fp.oe[0]) {
    T* dst = out + o0 + o1 * fp.oe[0] + plane * fp.oe[0] * fp.oe[1];
    const int nvalid = o1 + RB <= fp.oe[1] ? RB : (o1 < fp.oe[1] ? (int)(fp.oe[1] - o1) : 0);
#pragma unroll
    for (int o = 0; o < RB; ++o, dst += fp.oe[0])
      if (o < nvalid) *dst = acc[o];
  }
}
// f32 specialisation with packed arithmetic. Each thread owns TWO strips of RB2 outputs along dim 1, 32 columns apart in
// dim 0 (A = column lx, B = column lx + 32), and processes them as f32x2 lanes: per staged value pair (tA, tB) and tap weight w
//     FMUL2  p   = (w, w) * (tA, tB)          one rounding per lane, like the host's multiply
//     FFMA2  acc = acc * (1, 1) + p           exact times-one, then ONE rounding: the host's add
// i.e. one instruction per tap and output pair instead of the scalar kernel's FMUL + FADD per tap and output (r04: the 5x5 f32
// filter was FP32-issue-bound at 200 FMUL + 200 FADD per 8 outputs). Pairing across the two strips (instead of adjacent
// outputs) means a staged value always meets the same partner, so ptxas allocates each (tA, tB) in an aligned register pair
// straight from the two shared-memory loads: no MOVs. Tap order per output is unchanged (k1 ascending, k0 ascending inside:
// the host's column-major kernel walk, imfilter.rs:731-745), so the result stays bit-identical to the reference loop.
constexpr int RB2 = 4;  // 8 strips x 4 rows = the CTA's 32 output rows
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
template <int K0, int K1>
__global__ void __launch_bounds__(256)
imfilter_packed_f32_kernel(const float* __restrict__ img, const float* __restrict__ ker, float* __restrict__ out, const __grid_constant__ FilterParams fp) {
  constexpr int SX = FX + K0 - 1, SY = FY + K1 - 1;
  __shared__ float tile[SY][SX];
  __shared__ int map0[SX], map1[SY];
  const int64_t t0 = (int64_t)blockIdx.x * FX, t1 = (int64_t)blockIdx.y * FY;
  const uint64_t plane = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto remap = [&](int64_t c, int64_t len) -> int {
    if (c >= 0 && c < len) return (int)c;
    if (fp.padding == 0) return -1;
    return (int)(fp.padding == 1 ? clamp_index(c, len) : (fp.padding == 3 ? wrap_index(c, len) : reflect_index(c, len)));
  };
  const int64_t h0 = t0 + fp.base[0] - fp.origin[0], h1 = t1 + fp.base[1] - fp.origin[1];
  const bool interior = h0 >= 0 && h1 >= 0 && h0 + SX <= (int64_t)fp.ie[0] && h1 + SY <= (int64_t)fp.ie[1];
  if (!interior) {
    for (int i = tid; i < SX; i += 256) map0[i] = remap(h0 + i, (int64_t)fp.ie[0]);
    for (int i = tid; i < SY; i += 256) map1[i] = remap(h1 + i, (int64_t)fp.ie[1]);
  }
  float w[K0 * K1];  // application order (already flipped for convolution)
#pragma unroll
  for (int i = 0; i < K0 * K1; ++i) {
    const int k0 = i % K0, k1 = i / K0;
    w[i] = __ldg(ker + (fp.mode == 0 ? k0 + k1 * K0 : (K0 - 1 - k0) + (K1 - 1 - k1) * K0));
  }
  const float* src = img + plane * fp.ie[0] * fp.ie[1];
  if (interior) {
    const float* col = src + (uint64_t)h0 + (uint64_t)(h1 + warp) * fp.ie[0];
    const uint64_t step = 8 * fp.ie[0];
    for (int sy = warp; sy < SY; sy += 8, col += step) {
#pragma unroll
      for (int sx = lane; sx < SX; sx += 32) tile[sy][sx] = col[sx];
    }
  } else {
    __syncthreads();
    for (int sy = warp; sy < SY; sy += 8) {
      const int c = map1[sy];
      for (int sx = lane; sx < SX; sx += 32) {
        const int r = map0[sx];
        tile[sy][sx] = (r < 0 || c < 0) ? (float)fp.cval : src[(uint64_t)r + (uint64_t)c * fp.ie[0]];
      }
    }
  }
  __syncthreads();
  const int ly0 = warp * RB2;
  const unsigned long long one2 = pack_f32x2(fp.one, fp.one);
  unsigned long long acc[RB2];
#pragma unroll
  for (int o = 0; o < RB2; ++o) acc[o] = pack_f32x2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < RB2 + K1 - 1; ++j) {
#pragma unroll
    for (int k0 = 0; k0 < K0; ++k0) {
      const unsigned long long v = pack_f32x2(tile[ly0 + j][lane + k0], tile[ly0 + j][lane + 32 + k0]);
#pragma unroll
      for (int k1 = 0; k1 < K1; ++k1) {
        const int o = j - k1;
        if (o >= 0 && o < RB2) {
          const float wk = w[k0 + k1 * K0];
          unsigned long long prod;
          asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(prod) : "l"(pack_f32x2(wk, wk)), "l"(v));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[o]) : "l"(acc[o]), "l"(one2), "l"(prod));
        }
      }
    }
  }
  const uint64_t oA = (uint64_t)t0 + lane, oB = oA + 32, o1 = (uint64_t)t1 + ly0;
  float* dst = out + o1 * fp.oe[0] + plane * fp.oe[0] * fp.oe[1];
  const int nvalid = o1 + RB2 <= fp.oe[1] ? RB2 : (o1 < fp.oe[1] ? (int)(fp.oe[1] - o1) : 0);
#pragma unroll
  for (int o = 0; o < RB2; ++o, dst += fp.oe[0]) {
    if (o < nvalid) {
      float a, b;
      unpack_f32x2(acc[o], a, b);
      if (oA < fp.oe[0]) dst[oA] = a;
      if (oB < fp.oe[0]) dst[oB] = b;
    }
  }
}

// ---- TMA-staged persistent variant of the packed f32 kernel ------------------------------------------------------------------
// r06/r07: the packed kernel above needs 200 FMUL2/FFMA2 + 80 LDS per 8 outputs but runs the 2160x3840x3 frame in 111 us: a CTA's
// life is load tile (dependent LDG -> STS round trips) -> barrier -> math -> store, and the five CTAs an SM holds spend most of it
// waiting. Here a CTA is persistent (tiles interleaved across CTAs, so the tiles in flight are neighbours and share halos in L2)
// and warp-specialised:
//   * warp 8 (one lane) is the producer: for every tile of the CTA it waits until the 8 compute warps have released the ring
//     stage (mbarrier `empty`, count 8) and fetches tile + halo with ONE cp.async.bulk.tensor.3d (TMA, SASS UTMALDG) that
//     completes on the stage's `full` mbarrier: no load instructions, registers or warp waits on the data path;
//   * warps 0-7 each own RBW output rows x 64 columns of the tile: wait `full`, 128-bit-free LDS + packed math exactly as in the
//     kernel above, release the stage, store. There is no CTA-wide barrier on the interior path, so the warps drift apart and the
//     math of one overlaps the stores and waits of another (r09, first version with a __syncthreads per tile: barrier stalls 1.6
//     per issue, 88 us).
// Border tiles (halo outside the image: replicate / symmetric / circular / constant padding) are filled by the compute warps
// themselves through the index remap; they are the outer ring only. The TMA wants the box's inner start coordinate on a 16-byte
// boundary (scripts/imf_dev harness: "illegal instruction" at coordinate 126, fine at 124 / 128), so the fetch starts at the
// halo's first column rounded DOWN to a multiple of 4 floats (`shift` = 0..3 extra columns on the left). Arithmetic and tap
// order are exactly the packed kernel's, so the result is bit-identical to the reference loop (imfilter.rs:731-745).
constexpr int IMF_STAGES = 3;
template <int K, int RBW>
struct ImfTma {
  static constexpr int TY = 8 * RBW;  // output rows (dim 1) per tile: 8 compute warps x RBW
  static constexpr int SX = FX + K - 1, SY = TY + K - 1;
  static constexpr int SXP = (SX + 3 + 3) & ~3;  // + up to 3 shift columns, rows are whole 16-byte units
  static constexpr uint32_t BOX_BYTES = (uint32_t)SY * SXP * 4;
  static constexpr uint32_t STAGE_BYTES = (BOX_BYTES + 127) / 128 * 128;
  static constexpr uint32_t SMEM = IMF_STAGES * STAGE_BYTES + 128;
};
__device__ __forceinline__ uint32_t imf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool imf_mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(imf_smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) { atomicExch(err, 1); return false; }  // a protocol bug must surface as an error flag, never as a hung GPU
  }
}
// Producer flavour: lets the hardware park the lane for up to ~1 us per try instead of re-issuing the test (r53 ncu: the lone
// producer lane executed 14 try_wait round trips of 9 instructions per tile-warp, 12 % of the kernel's issue slots).
__device__ __forceinline__ bool imf_mbar_wait_parked(uint64_t* bar, uint32_t parity, int* err) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(imf_smem_u32(bar)), "r"(parity), "r"(1000u) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) { atomicExch(err, 1); return false; }
  }
}
// Tile t = blockIdx.x + it * gridDim.x as (tile column, tile row, plane), advanced incrementally: the stride is decomposed once,
// a step is three adds and two compare-subtracts. (r53 ncu/SASS: decomposing t with two 32-bit divisions per tile, in 64-bit
// coordinates, cost ~100 dependent instructions per tile per warp -- I2F/MUFU.RCP/F2I chains the warp sits out -- next to the
// 400 packed math instructions of a 5x5 tile.)
// ---- imf_tile_iter begin (tests/test_host_logic.py compiles this struct for the host and checks it against the divisions)
struct ImfTileIter {
  uint32_t tx, ty, plane, dx, dy, dp, ntx, nty;
  __device__ __forceinline__ void init(uint32_t first, uint32_t stride, uint32_t ntx_, uint32_t nty_) {
    ntx = ntx_; nty = nty_;
    tx = first % ntx; uint32_t r = first / ntx; ty = r % nty; plane = r / nty;
    dx = stride % ntx; r = stride / ntx; dy = r % nty; dp = r / nty;
  }
  __device__ __forceinline__ void next() {
    tx += dx;
    uint32_t c = 0;
    if (tx >= ntx) { tx -= ntx; c = 1; }
    ty += dy + c;
    plane += dp;
    if (ty >= nty) { ty -= nty; ++plane; }
  }
};
// ---- imf_tile_iter end
// MODE 0: scalar FMUL + FADD; 1: packed FMUL2 + FFMA2(x1); 2: mixed -- the first half of a warp's rows packed, the second half scalar
// Resident CTAs the register allocation is held to: the K*K weights live in registers, so 7x7 gets two CTAs per SM; up to 5x5
// four fit (RBW = 4: 3 x 10 KB stages each) or three (RBW = 8).
// MB = resident CTAs the register allocation is held to. 5x5 at MB = 3 (75 registers) leaves ptxas ONE product temporary, so every
// FMUL2 is followed at once by the FFMA2 that consumes it; at MB = 2 (91 registers used) products are issued 2-4 instructions
// ahead of their accumulation. Measured (r56, 2160x3840x3): 63.6 us at MB = 3 vs 66.4 us at MB = 2 -- the third CTA's eight warps
// hide that latency better than the wider schedule does, so MB = 3 stays the default (RUNMAT_B200_IMFILTER_MINB2 = the other).
template <int K, int RBW, int MODE, int MB>
__global__ void __launch_bounds__(288, MB)
imfilter_tma_f32_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ img, const float* __restrict__ ker, float* __restrict__ out,
                        const __grid_constant__ FilterParams fp, uint32_t ntx, uint32_t nty, uint32_t ntiles, int* __restrict__ err,
                        uint32_t num_sms, uint32_t cta_stagger_ns, uint32_t warp_stagger_ns) {
  using C = ImfTma<K, RBW>;
  constexpr int SX = C::SX, SY = C::SY, SXP = C::SXP;
  extern __shared__ __align__(128) unsigned char imf_dsm_raw[];
  // 128-byte alignment of the stages by an OFFSET on the shared array (a pointer round trip through an integer would turn the
  // tile reads into generic LD instead of LDS)
  unsigned char* dsm = imf_dsm_raw + ((128u - (imf_smem_u32(imf_dsm_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t full[IMF_STAGES], empty[IMF_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < IMF_STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(imf_smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(imf_smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_prologue();  // barrier set-up above overlaps the previous kernel's drain; no global access before this point
  struct TileAt { int t0, t1, h0, h1; uint32_t plane; bool interior, tma; };  // image extents and tile counts are below 2^31 (launch_imfilter_tma_rb)
  // columns the fetch starts left of the halo; the same for every tile of a launch (tile origins are multiples of FX = 64)
  const int bo0 = (int)(fp.base[0] - fp.origin[0]), bo1 = (int)(fp.base[1] - fp.origin[1]);
  const int shift = ((bo0 % 4) + 4) % 4;
  const int ie0 = (int)fp.ie[0], ie1 = (int)fp.ie[1];
  const int padding = fp.padding;
  auto tile_at = [&](const ImfTileIter& ti) {
    TileAt a;
    a.t0 = (int)ti.tx * FX;
    a.t1 = (int)ti.ty * C::TY;
    a.plane = ti.plane;
    a.h0 = a.t0 + bo0;
    a.h1 = a.t1 + bo1;
    a.interior = a.h0 >= 0 && a.h1 >= 0 && a.h0 + SX <= ie0 && a.h1 + SY <= ie1;
    // Border tiles are fetched by the TMA as well (out-of-image elements arrive as zeros) and patched in shared memory from the
    // in-image part of the same box; only circular padding reads from the far side of the image and keeps the manual fill.
    a.tma = a.interior || padding != 3;
    return a;
  };
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ... (the grid never exceeds the tile count)
  const uint32_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  ImfTileIter ti;
  ti.init(blockIdx.x, gridDim.x, ntx, nty);
  if (warp == 8) {
    // ---- producer ------------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t round = 0;  // it / IMF_STAGES
      for (uint32_t it = 0; it < my_tiles; ++it, ti.next(), s = (s + 1 == IMF_STAGES ? 0 : s + 1), round += (s == 0)) {
        const TileAt a = tile_at(ti);
        if (!a.tma) continue;  // circular-padding border tiles are filled by the compute warps (they wait for `empty` themselves)
        if (it >= IMF_STAGES && !imf_mbar_wait_parked(&empty[s], (round - 1) & 1u, err)) break;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(imf_smem_u32(&full[s])), "r"(C::BOX_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(imf_smem_u32(dsm + (size_t)s * C::STAGE_BYTES)), "l"(&tm), "r"(imf_smem_u32(&full[s])), "r"(a.h0 - shift), "r"(a.h1), "r"((int)a.plane)
                     : "memory");
      }
    }
    return;
  }
  // ---- compute warps -----------------------------------------------------------------------------------------------------------
  // Phase stagger. All CTAs start together and every tile costs the same, so without help the warps of an SM sit in their
  // non-math phases (barrier wait, tile bookkeeping, stores) at the same time and the FP32 pipe idles (r53 ncu: no eligible warp
  // in 36 % of the cycles while the average is 1.9 eligible warps). The k-th co-resident CTA (blockIdx.x / #SMs) starts its
  // compute warps k * cta_stagger_ns late, the upper half of each CTA's warps another warp_stagger_ns; the producer is not
  // delayed, so the ring is full when they start, and the offsets persist because tile times are equal.
  {
    const uint32_t delay = (blockIdx.x / num_sms) * cta_stagger_ns + (warp >= 4 ? warp_stagger_ns : 0u);
    if (delay) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do { __nanosleep(64); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < delay);
    }
  }
  float w[K * K];  // application order (already flipped for convolution)
#pragma unroll
  for (int i = 0; i < K * K; ++i) {
    const int k0 = i % K, k1 = i / K;
    w[i] = __ldg(ker + (fp.mode == 0 ? k0 + k1 * K : (K - 1 - k0) + (K - 1 - k1) * K));
  }
  auto remap = [&](int64_t c, int64_t len) -> int {
    if (c >= 0 && c < len) return (int)c;
    if (fp.padding == 0) return -1;
    return (int)(fp.padding == 1 ? clamp_index(c, len) : (fp.padding == 3 ? wrap_index(c, len) : reflect_index(c, len)));
  };
  const int ly0 = warp * RBW;
  const unsigned long long one2 = pack_f32x2(fp.one, fp.one);
  uint32_t phase = 0;  // bit s: parity of the next completion of full[s] (it advances on TMA-fetched tiles only)
  int s = 0;
  uint32_t round = 0;  // it / IMF_STAGES
  for (uint32_t it = 0; it < my_tiles; ++it, ti.next(), s = (s + 1 == IMF_STAGES ? 0 : s + 1), round += (s == 0)) {
    float* tile = (float*)(dsm + (size_t)s * C::STAGE_BYTES);
    const TileAt a = tile_at(ti);
    if (a.tma) {
      imf_mbar_wait(&full[s], (phase >> s) & 1u, err);
      phase ^= 1u << s;
      if (!a.interior) {
        // patch the out-of-image elements of the box (zeros from the TMA): constant -> cval; replicate / symmetric -> the value
        // at the remapped coordinate, which lies in the in-image part of this same box (read from global memory if it does not)
        // Only the out-of-image cells are visited (r53: scanning all SY*SX cells with 64-bit index math made a border tile cost
        // ~3x an interior one: 15 % of the compute warps' samples for 9 % of the tiles): the nl / nr out-of-image columns of every
        // row first, then the nt / nb out-of-image rows of the columns in between.
        const float* src = img + (uint64_t)a.plane * fp.ie[0] * fp.ie[1];
        int nl = a.h0 < 0 ? -a.h0 : 0, nr = a.h0 + SX > ie0 ? a.h0 + SX - ie0 : 0;
        int nt = a.h1 < 0 ? -a.h1 : 0, nb = a.h1 + SY > ie1 ? a.h1 + SY - ie1 : 0;
        nl = nl < SX ? nl : SX; nr = nr < SX - nl ? nr : SX - nl;
        nt = nt < SY ? nt : SY; nb = nb < SY - nt ? nb : SY - nt;
        const int wside = nl + nr, wmid = SX - wside, nside = SY * wside, ncell = nside + (nt + nb) * wmid;
        for (int idx = tid; idx < ncell; idx += 256) {
          int sx, sy;
          if (idx < nside) {
            sy = idx / wside;
            const int c = idx - sy * wside;
            sx = c < nl ? c : SX - nr + (c - nl);
          } else {
            const int j = idx - nside, rr = j / wmid;
            sx = nl + (j - rr * wmid);
            sy = rr < nt ? rr : SY - nb + (rr - nt);
          }
          const int gx = a.h0 + sx, gy = a.h1 + sy;
          const int r = remap(gx, (int64_t)ie0), c = remap(gy, (int64_t)ie1);
          float v;
          if (r < 0 || c < 0) v = (float)fp.cval;
          else {
            const int bx = r - a.h0, by = c - a.h1;
            v = (bx >= 0 && bx < SX && by >= 0 && by < SY) ? tile[by * SXP + shift + bx] : src[(uint64_t)r + (uint64_t)c * fp.ie[0]];
          }
          tile[sy * SXP + shift + sx] = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    } else {
      // every compute warp has released the stage's previous use; then fill it cooperatively and meet on a named barrier
      if (it >= IMF_STAGES) imf_mbar_wait(&empty[s], (round - 1) & 1u, err);
      const float* src = img + (uint64_t)a.plane * fp.ie[0] * fp.ie[1];
      for (int sy = warp; sy < SY; sy += 8) {
        const int c = remap(a.h1 + sy, (int64_t)fp.ie[1]);
        for (int sx = lane; sx < SX; sx += 32) {
          const int r = remap(a.h0 + sx, (int64_t)fp.ie[0]);
          tile[sy * SXP + shift + sx] = (r < 0 || c < 0) ? (float)fp.cval : src[(uint64_t)r + (uint64_t)c * fp.ie[0]];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    unsigned long long acc[RBW];
    float accA[RBW], accB[RBW];
#pragma unroll
    for (int o = 0; o < RBW; ++o) { acc[o] = pack_f32x2(0.0f, 0.0f); accA[o] = 0.0f; accB[o] = 0.0f; }
#pragma unroll
    for (int j = 0; j < RBW + K - 1; ++j) {
#pragma unroll
      for (int k0 = 0; k0 < K; ++k0) {
        const float tA = tile[(ly0 + j) * SXP + shift + lane + k0], tB = tile[(ly0 + j) * SXP + shift + lane + 32 + k0];
        const unsigned long long v = pack_f32x2(tA, tB);
        // The K products of one loaded pair first, then the K accumulations (K different outputs). r53 SASS: with the multiply
        // and its add written back to back the scheduler walked ONE output's chain at a time -- five dependent FFMA2 in a row, a
        // warp issued one instruction per 4.7 cycles and it took ~5 warps per sub-partition in the loop to fill the FP32 pipe.
        // `volatile` pins this order; each output still receives its taps in the host's order.
        unsigned long long prod[K];
        if (MODE == 1) {
#pragma unroll
          for (int k1 = 0; k1 < K; ++k1) {
            const int o = j - k1;
            if (o >= 0 && o < RBW) {
              const float wk = w[k0 + k1 * K];
              asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(prod[k1]) : "l"(pack_f32x2(wk, wk)), "l"(v));
            }
          }
#pragma unroll
          for (int k1 = 0; k1 < K; ++k1) {
            const int o = j - k1;
            if (o >= 0 && o < RBW) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[o]) : "l"(acc[o]), "l"(one2), "l"(prod[k1]));
          }
        }
#pragma unroll
        for (int k1 = 0; k1 < K; ++k1) {
          const int o = j - k1;
          if (MODE != 1 && o >= 0 && o < RBW) {
            const float wk = w[k0 + k1 * K];
            if (MODE == 2 && o < RBW / 2) {  // o is a compile-time constant after unrolling
              unsigned long long pr;
              asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pack_f32x2(wk, wk)), "l"(v));
              asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[o]) : "l"(acc[o]), "l"(one2), "l"(pr));
            } else {
              // scalar FMUL + FADD (the library is compiled -fmad=false): one rounding after the multiply, one after the add
              accA[o] = __fadd_rn(accA[o], __fmul_rn(wk, tA));
              accB[o] = __fadd_rn(accB[o], __fmul_rn(wk, tB));
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(imf_smem_u32(&empty[s])) : "memory");  // this warp's reads of the stage are done
    // stores: one 64-bit base per tile, then a pointer step per row; tiles that lie wholly inside the output (all but the last
    // tile row / column) take the unguarded path (r53: the guarded form with per-store 64-bit compares was 165 instructions)
    const int oe0 = (int)fp.oe[0], oe1 = (int)fp.oe[1];
    const int oA = a.t0 + lane, o1 = a.t1 + ly0;
    float* dst = out + ((uint64_t)a.plane * (uint32_t)oe1 + (uint32_t)o1) * (uint64_t)(uint32_t)oe0 + (uint32_t)oA;
    if (a.t0 + FX <= oe0 && o1 + RBW <= oe1) {
#pragma unroll
      for (int o = 0; o < RBW; ++o, dst += oe0) {
        float va, vb;
        if (MODE == 1 || (MODE == 2 && o < RBW / 2)) unpack_f32x2(acc[o], va, vb);
        else { va = accA[o]; vb = accB[o]; }
        __stcs(dst, va);
        __stcs(dst + 32, vb);
      }
    } else {
      const int nvalid = o1 + RBW <= oe1 ? RBW : (o1 < oe1 ? oe1 - o1 : 0);
#pragma unroll
      for (int o = 0; o < RBW; ++o, dst += oe0) {
        if (o < nvalid) {
          float va, vb;
          if (MODE == 1 || (MODE == 2 && o < RBW / 2)) unpack_f32x2(acc[o], va, vb);
          else { va = accA[o]; vb = accB[o]; }
          if (oA < oe0) __stcs(dst, va);
          if (oA + 32 < oe0) __stcs(dst + 32, vb);
        }
      }
    }
  }
}

typedef CUresult (*ImfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static ImfEncodeTiledFn imf_encode_tiled() {
  static ImfEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (ImfEncodeTiledFn)ptr;
    cudaGetLastError();
  });
  return fn;
}
// Returns false when the TMA path does not apply (the caller then launches the per-tile kernel).
template <int K, int RBW, int MODE, int MB = (K <= 5 ? (RBW == 4 ? 4 : 3) : 2)>
static bool launch_imfilter_tma_rb(rm_provider* p, const float* a, const float* k, float* o, const FilterParams& fp) {
  using C = ImfTma<K, RBW>;
  const uint64_t ntx = (fp.oe[0] + FX - 1) / FX, nty = (fp.oe[1] + C::TY - 1) / C::TY, ntiles = ntx * nty * fp.oe[2];
  // global strides must be multiples of 16 bytes; coordinates are 32-bit; small problems keep the one-CTA-per-tile kernel
  if (fp.ie[0] % 4 != 0 || fp.ie[0] >= (1ull << 31) || fp.ie[1] >= (1ull << 31) || fp.ie[2] >= (1ull << 31) || ntiles >= (1ull << 31) ||
      ntiles < (uint64_t)p->prop.multiProcessorCount * 8 || ((uintptr_t)a & 15) != 0)
    return false;
  ImfEncodeTiledFn enc = imf_encode_tiled();
  if (!enc) return false;
  if (!p->dev_flags) {
    std::lock_guard<std::mutex> lk(p->scratch_mu);
    if (!p->dev_flags) {
      int* f = nullptr;
      if (cudaMalloc((void**)&f, 64 * sizeof(int)) != cudaSuccess || cudaMemset(f, 0, 64 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return false; }
      p->dev_flags = f;
    }
  }
  CUtensorMap tm;
  cuuint64_t dims[3] = {fp.ie[0], fp.ie[1], fp.ie[2]};
  cuuint64_t strides[2] = {fp.ie[0] * 4, fp.ie[0] * fp.ie[1] * 4};
  cuuint32_t box[3] = {(cuuint32_t)C::SXP, (cuuint32_t)C::SY, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  static bool attr_set = false;  // > 48 KB of dynamic shared memory needs the opt-in (idempotent; a benign race)
  if (!attr_set) { cudaFuncSetAttribute(imfilter_tma_f32_kernel<K, RBW, MODE, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM); attr_set = true; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, imfilter_tma_f32_kernel<K, RBW, MODE, MB>, 288, C::SMEM) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return false; }
  if (const char* e = getenv("RUNMAT_B200_IMFILTER_CTAS")) { const int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }
  const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)p->prop.multiProcessorCount * per_sm);
  uint32_t cta_stagger = 0, warp_stagger = 0;
  if (const char* e = getenv("RUNMAT_B200_IMFILTER_STAGGER")) cta_stagger = (uint32_t)atoi(e);
  if (const char* e = getenv("RUNMAT_B200_IMFILTER_WSTAGGER")) warp_stagger = (uint32_t)atoi(e);
  // r58: programmatic dependent launch pays from 5x5 up (61.5 -> 58.3 us, 7x7 96.1 -> 93.5); the bandwidth-bound 3x3 loses 3 % to
  // the next launch's early-resident CTAs (42.2 -> 43.7 us), so it keeps the plain launch
  launch_pdl(p->launch_overlap && K > 3, imfilter_tma_f32_kernel<K, RBW, MODE, MB>, dim3(grid), dim3(288), C::SMEM, p->stream, tm, a, k, o, fp, (uint32_t)ntx, (uint32_t)nty,
             (uint32_t)ntiles, (int*)p->dev_flags + 0, (uint32_t)p->prop.multiProcessorCount, cta_stagger, warp_stagger);
  return true;
}
template <int K>
static bool launch_imfilter_tma(rm_provider* p, const float* a, const float* k, float* o, const FilterParams& fp) {
  if (getenv("RUNMAT_B200_IMFILTER_NO_TMA")) return false;
  int rbw = 8;  // output rows per compute warp (tile = 64 x 8*rbw)
  if (const char* e = getenv("RUNMAT_B200_IMFILTER_RBW")) rbw = atoi(e);
  int mode = 1;
  if (const char* e = getenv("RUNMAT_B200_IMFILTER_MODE")) mode = atoi(e);
  if constexpr (K == 5) {
    if (rbw == 8 && mode == 1 && getenv("RUNMAT_B200_IMFILTER_MINB2")) return launch_imfilter_tma_rb<K, 8, 1, 2>(p, a, k, o, fp);  // A/B: the 91-register build
  }
  if (rbw == 4) return mode == 0 ? launch_imfilter_tma_rb<K, 4, 0>(p, a, k, o, fp) : mode == 2 ? launch_imfilter_tma_rb<K, 4, 2>(p, a, k, o, fp) : launch_imfilter_tma_rb<K, 4, 1>(p, a, k, o, fp);
  return mode == 0 ? launch_imfilter_tma_rb<K, 8, 0>(p, a, k, o, fp) : mode == 2 ? launch_imfilter_tma_rb<K, 8, 2>(p, a, k, o, fp) : launch_imfilter_tma_rb<K, 8, 1>(p, a, k, o, fp);
}

template <typename T, int K>
static void launch_regblock_k(rm_provider* p, dim3 grid, const T* a, const T* k, T* o, const FilterParams& fp) {
  if constexpr (std::is_same<T, float>::value) {
    if (!getenv("RUNMAT_B200_IMFILTER_SCALAR")) {
      if (launch_imfilter_tma<K>(p, a, k, o, fp)) return;
      imfilter_packed_f32_kernel<K, K><<<grid, 256, 0, p->stream>>>(a, k, o, fp);
      return;
    }
  }
  imfilter_regblock_kernel<T, K, K><<<grid, 256, 0, p->stream>>>(a, k, o, fp);
}
template <typename T>
static bool launch_regblock(rm_provider* p, const void* pi, const void* pk, void* po, const FilterParams& fp) {
  if (fp.ke[2] != 1 || fp.ke[0] != fp.ke[1]) return false;
  dim3 grid((unsigned)((fp.oe[0] + FX - 1) / FX), (unsigned)((fp.oe[1] + FY - 1) / FY), (unsigned)fp.oe[2]);
  if (grid.y > 65535 || grid.z > 65535) return false;
  const T* a = (const T*)pi; const T* k = (const T*)pk; T* o = (T*)po;
  switch (fp.ke[0]) {
    case 3: launch_regblock_k<T, 3>(p, grid, a, k, o, fp); return true;
    case 5: launch_regblock_k<T, 5>(p, grid, a, k, o, fp); return true;
    case 7: launch_regblock_k<T, 7>(p, grid, a, k, o, fp); return true;
    default: return false;
  }
}

// ---- value + index reductions along one dim of [pre, n, post] (simple_provider.rs:7387-7445) ---------------------------
template <typename T, bool IS_MIN>
__global__ void minmax_dim_kernel(const T* __restrict__ a, T* __restrict__ vals, T* __restrict__ idx, uint64_t pre, uint64_t n, uint64_t post) {
  // one warp per slice; lanes stride over n, then a shuffle argmin/argmax with smallest-index tie-break
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= pre * post) return;
  const uint64_t s_lo = warp % pre, s_hi = warp / pre;
  const T* base = a + s_lo + s_hi * pre * n;
  T best = IS_MIN ? (T)INFINITY : (T)-INFINITY;
  uint64_t best_i = 0;  // host: idx initialised to 1 (first element) when nothing compares strictly better
  for (uint64_t r = lane; r < n; r += 32) {
    const T v = base[r * pre];
    if (IS_MIN ? v < best : v > best) { best = v; best_i = r; }
  }
  // lanes that saw no strictly-better value keep (identity, 0); ties resolve to the smallest index
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const T ov = __shfl_xor_sync(0xffffffffu, best, off);
    const uint64_t oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    const bool better = IS_MIN ? ov < best : ov > best;
    if (better || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane == 0) { vals[warp] = best; idx[warp] = (T)(best_i + 1); }
}

}  // namespace

}  // namespace rm

using namespace rm;

RM_EXPORT rm_status rm_image_normalize(rm_provider* p, const rm_handle* input, const rm_image_normalize_desc* d, rm_handle* out) {
  RM_REQUIRE(p && input && d && out, RM_INVALID_ARG, "image_normalize: bad arguments");
  // simple_provider.rs:7899-7917
  RM_REQUIRE(isfinite(d->epsilon), RM_ERROR, "image_normalize: epsilon must be finite");
  RM_REQUIRE(d->epsilon >= 0.0, RM_ERROR, "image_normalize: epsilon must be non-negative");
  RM_REQUIRE(input->rank == 3, RM_ERROR, "image_normalize: expected 3-D tensor, got rank %u", input->rank);
  RM_REQUIRE(input->shape[0] == d->batch && input->shape[1] == d->height && input->shape[2] == d->width, RM_ERROR,
             "image_normalize: descriptor dims (%llu, %llu, %llu) do not match tensor shape", (unsigned long long)d->batch,
             (unsigned long long)d->height, (unsigned long long)d->width);
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, input, &src, nullptr));
  const uint64_t B = d->batch, P = d->height * d->width, total = B * P;
  void* dst;
  RM_TRY(alloc_tensor(p, input->shape, 3, out, &dst));
  if (total == 0) return RM_OK;
  RM_REQUIRE(B <= 1024, RM_UNSUPPORTED, "image_normalize: batch %llu > 1024 not supported by provider", (unsigned long long)B);

  const int VEC = p->precision == RM_F64 ? 2 : 4;
  const uint32_t sms = (uint32_t)p->prop.multiProcessorCount;
  // fast path: lanes stay on a fixed image when the grid stride (in elements) is a multiple of B
  uint32_t fast_blocks = 0;
  if (B % VEC == 0 && total % VEC == 0) {
    for (uint32_t mult = 8; mult >= 1 && !fast_blocks; --mult) {
      const uint64_t blocks = (uint64_t)sms * mult;
      if ((blocks * 256ull * VEC) % B == 0 && blocks * 256ull * 2 <= total / VEC + blocks * 256ull) fast_blocks = (uint32_t)blocks;
    }
  }
  const uint32_t threads = 1024 - (1024 % (uint32_t)B);  // (generic path) whole number of pixel lanes
  const uint32_t lanes = threads / (uint32_t)B;
  uint32_t nblocks = fast_blocks ? fast_blocks : (uint32_t)std::min<uint64_t>((uint64_t)sms * 2, std::max<uint64_t>(1, P / (lanes * 4ull)));
  const size_t partial_bytes = (size_t)nblocks * 2 * B * sizeof(double);
  const size_t stats_off = ((partial_bytes + 255) / 256) * 256;
  std::lock_guard<std::mutex> scratch_lock(p->scratch_mu);  // held until the kernels that use the scratch are enqueued
  rm_status st = ensure_scratch(p, stats_off + 2 * B * sizeof(double));
  if (st != RM_OK) { rm_free(p, out); return st; }
  double* partial = (double*)p->reduce_scratch;
  double* stats = (double*)((char*)p->reduce_scratch + stats_off);
  const size_t sh = fast_blocks ? (size_t)2 * B * sizeof(double) : (size_t)lanes * 2 * B * sizeof(double);
  NormParams np{d->has_gain, d->has_bias, d->has_gamma, d->clamp_zero, d->gain, d->bias, d->gamma};
  const unsigned ngrid = (unsigned)std::min<uint64_t>((total / 4 + 255) / 256 + 1, (uint64_t)sms * 16);
  const uint64_t groups = B / (uint64_t)VEC;  // image groups a warp's lanes cycle through on the fast path
  const bool fold = fast_blocks && B % VEC == 0 && groups >= 1 && groups <= 32 && (groups & (groups - 1)) == 0 && !getenv("RUNMAT_B200_MOMENTS_ATOMIC");
  const size_t sh_fold = (size_t)8 * 2 * B * sizeof(double);
  // end-to-front normalise sweep over a tensor the moments sweep has just read front-to-end (L2 reuse of its tail); A/B switches
  const int rev = getenv("RUNMAT_B200_NORMALIZE_FORWARD") ? 0 : 1;
  const bool pdl = p->launch_overlap;  // programmatic dependent launch of the three-kernel chain (common.h launch_pdl)
  const bool keep = rev && getenv("RUNMAT_B200_MOMENTS_KEEP");  // r51: plain loads measured slower (0.1473 vs 0.1421 ms), evict-first stays
  if (p->precision == RM_F64) {
    if (fold) {
      launch_pdl(pdl, keep ? moments_partial_fold_kernel<double, 2, true> : moments_partial_fold_kernel<double, 2, false>, dim3(nblocks), dim3(256), sh_fold, p->stream,
                 (const double*)src, (uint32_t)B, total, partial);
    } else if (fast_blocks) {
      moments_partial_vec_kernel<double, 2><<<nblocks, 256, sh, p->stream>>>((const double*)src, (uint32_t)B, total, partial);
    } else {
      cudaFuncSetAttribute(moments_partial_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
      moments_partial_kernel<double><<<nblocks, threads, sh, p->stream>>>((const double*)src, B, P, partial);
    }
    launch_pdl(pdl, moments_finalize_kernel<double>, dim3((unsigned)B), dim3(128), 0, p->stream, (const double*)src, (const double*)partial, nblocks, B, P, d->epsilon, stats);
    if (fast_blocks) launch_pdl(pdl, normalize_fixed_kernel<double, 2>, dim3(fast_blocks), dim3(256), 0, p->stream, (const double*)src, (double*)dst, (uint32_t)B, total, (const double*)stats, np, rev);
    else normalize_kernel<double, 2><<<ngrid, 256, 0, p->stream>>>((const double*)src, (double*)dst, B, total, stats, np);
  } else {
    if (fold) {
      launch_pdl(pdl, keep ? moments_partial_fold_kernel<float, 4, true> : moments_partial_fold_kernel<float, 4, false>, dim3(nblocks), dim3(256), sh_fold, p->stream,
                 (const float*)src, (uint32_t)B, total, partial);
    } else if (fast_blocks) {
      moments_partial_vec_kernel<float, 4><<<nblocks, 256, sh, p->stream>>>((const float*)src, (uint32_t)B, total, partial);
    } else {
      cudaFuncSetAttribute(moments_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
      moments_partial_kernel<float><<<nblocks, threads, sh, p->stream>>>((const float*)src, B, P, partial);
    }
    launch_pdl(pdl, moments_finalize_kernel<float>, dim3((unsigned)B), dim3(128), 0, p->stream, (const float*)src, (const double*)partial, nblocks, B, P, d->epsilon, stats);
    if (fast_blocks && d->clamp_zero && d->has_gamma && (float)d->gamma != 0.0f && !getenv("RUNMAT_B200_NORMALIZE_GENERIC")) {
      // 256-bit accesses when a thread's 8 lanes stay on fixed images (B % 8 == 0 and the grid stride a multiple of B)
      if (B % 8 == 0 && total % 8 == 0 && ((uint64_t)fast_blocks * 256ull * 8) % B == 0 && !getenv("RUNMAT_B200_NORMALIZE_VEC4"))
        launch_pdl(pdl, normalize_fixed_clampgamma_f32_kernel<8>, dim3(fast_blocks), dim3(256), 0, p->stream, (const float*)src, (float*)dst, (uint32_t)B, total, (const double*)stats, np, rev);
      else
        launch_pdl(pdl, normalize_fixed_clampgamma_f32_kernel<4>, dim3(fast_blocks), dim3(256), 0, p->stream, (const float*)src, (float*)dst, (uint32_t)B, total, (const double*)stats, np, rev);
    }
    else if (fast_blocks) launch_pdl(pdl, normalize_fixed_kernel<float, 4>, dim3(fast_blocks), dim3(256), 0, p->stream, (const float*)src, (float*)dst, (uint32_t)B, total, (const double*)stats, np, rev);
    else normalize_kernel<float, 4><<<ngrid, 256, 0, p->stream>>>((const float*)src, (float*)dst, B, total, stats, np);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rm_free(p, out); return fail(RM_ERROR, "image_normalize launch failed: %s", cudaGetErrorString(e)); }
  count_launch(p, 3);
  return RM_OK;
}

// Launches the tiled (2-D kernel) or generic (N-D kernel) filter kernel for a prepared FilterParams.
static void launch_filter(rm_provider* p, const void* pi, const void* pk, void* po, uint64_t total, const FilterParams& fp) {
  const bool tiled = fp.ke[2] == 1 && fp.ke[0] <= IMF_MAXK && fp.ke[1] <= IMF_MAXK && fp.oe[2] <= 65535 && (fp.oe[1] + IMF_TY - 1) / IMF_TY <= 65535 &&
                     fp.ie[0] < (1ull << 31) && fp.ie[1] < (1ull << 31);
  const bool small_dims = fp.ie[0] < (1ull << 31) && fp.ie[1] < (1ull << 31);
  if (small_dims && !getenv("RUNMAT_B200_IMFILTER_GENERIC") &&
      (p->precision == RM_F64 ? launch_regblock<double>(p, pi, pk, po, fp) : launch_regblock<float>(p, pi, pk, po, fp))) {
    // register-blocked 3x3 / 5x5 / 7x7 path
  } else if (tiled) {
    const int SX = IMF_TX + (int)fp.ke[0] - 1, SY = IMF_TY + (int)fp.ke[1] - 1;
    const size_t sh = (size_t)(SX * SY + (int)(fp.ke[0] * fp.ke[1])) * p->elem_size() + (size_t)(SX + SY) * sizeof(int) + 16;
    dim3 grid((unsigned)((fp.oe[0] + IMF_TX - 1) / IMF_TX), (unsigned)((fp.oe[1] + IMF_TY - 1) / IMF_TY), (unsigned)fp.oe[2]);
    if (p->precision == RM_F64) imfilter_tiled_kernel<double><<<grid, 256, sh, p->stream>>>((const double*)pi, (const double*)pk, (double*)po, fp);
    else imfilter_tiled_kernel<float><<<grid, 256, sh, p->stream>>>((const float*)pi, (const float*)pk, (float*)po, fp);
  } else {
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)p->prop.multiProcessorCount * 32);
    if (p->precision == RM_F64) imfilter_kernel<double><<<grid, 256, 0, p->stream>>>((const double*)pi, (const double*)pk, (double*)po, total, fp);
    else imfilter_kernel<float><<<grid, 256, 0, p->stream>>>((const float*)pi, (const float*)pk, (float*)po, total, fp);
  }
}

RM_EXPORT rm_status rm_imfilter(rm_provider* p, const rm_handle* image, const rm_handle* kernel, const rm_imfilter_options* opt, rm_handle* out) {
  RM_REQUIRE(p && image && kernel && opt && out, RM_INVALID_ARG, "imfilter: bad arguments");
  DeviceGuard g(p->ordinal);
  void *pi, *pk;
  uint64_t ni, nk;
  RM_TRY(resolve(p, image, &pi, &ni));
  RM_TRY(resolve(p, kernel, &pk, &nk));
  RM_REQUIRE(nk > 0, RM_ERROR, "imfilter: filter must be non-empty along every dimension");  // imfilter.rs:483-488
  const uint32_t rank = std::max<uint32_t>(std::max(image->rank, kernel->rank), 2);
  RM_REQUIRE(rank <= 3, RM_UNSUPPORTED, "imfilter: rank %u not supported by provider (max 3)", rank);
  FilterParams fp{};
  for (int d = 0; d < 3; ++d) {
    fp.ie[d] = d < (int)image->rank ? image->shape[d] : 1;
    fp.ke[d] = d < (int)kernel->rank ? kernel->shape[d] : 1;
    RM_REQUIRE(!(d >= (int)std::max<uint32_t>(image->rank, 2) && fp.ke[d] > 1), RM_ERROR,
               "imfilter: filter dimension %d is %llu, but the image has no corresponding axis", d + 1, (unsigned long long)fp.ke[d]);
    RM_REQUIRE(fp.ie[d] != 0, RM_ERROR, "imfilter: image must not have zero-length dimensions");
    fp.origin[d] = (int64_t)(fp.ke[d] / 2);
    if (opt->shape == RM_IMF_FULL) { fp.oe[d] = fp.ie[d] + fp.ke[d] - 1; fp.base[d] = fp.origin[d] - ((int64_t)fp.ke[d] - 1); }
    else if (opt->shape == RM_IMF_SAME) { fp.oe[d] = fp.ie[d]; fp.base[d] = 0; }
    else { fp.oe[d] = fp.ie[d] >= fp.ke[d] ? fp.ie[d] - fp.ke[d] + 1 : 0; fp.base[d] = fp.origin[d]; }
  }
  fp.padding = (int)opt->padding;
  fp.mode = (int)opt->mode;
  fp.cval = opt->constant_value;
  // final_shape: trailing singleton dims beyond the image rank are dropped (imfilter.rs:541-547)
  uint64_t oshape[3] = {fp.oe[0], fp.oe[1], fp.oe[2]};
  uint32_t orank = 3;
  while (orank > std::max<uint32_t>(image->rank, 2) && oshape[orank - 1] == 1) --orank;
  void* po;
  RM_TRY(alloc_tensor(p, oshape, orank, out, &po));
  const uint64_t total = fp.oe[0] * fp.oe[1] * fp.oe[2];
  if (total == 0) return RM_OK;
  launch_filter(p, pi, pk, po, total, fp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rm_free(p, out); return fail(RM_ERROR, "imfilter launch failed: %s", cudaGetErrorString(e)); }
  count_launch(p);
  return RM_OK;
}

RM_EXPORT rm_status rm_conv2d(rm_provider* p, const rm_handle* signal, const rm_handle* kernel, rm_conv_mode mode, rm_handle* out) {
  RM_REQUIRE(p && signal && kernel && out, RM_INVALID_ARG, "conv2d: bad arguments");
  RM_REQUIRE(signal->rank <= 2 && kernel->rank <= 2, RM_ERROR, "conv2d: inputs must be 2-D matrices");
  DeviceGuard g(p->ordinal);
  void *pi, *pk;
  uint64_t ni, nk;
  RM_TRY(resolve(p, signal, &pi, &ni));
  RM_TRY(resolve(p, kernel, &pk, &nk));
  const uint64_t sr = signal->rank >= 1 ? signal->shape[0] : 1, sc = signal->rank >= 2 ? signal->shape[1] : 1;
  const uint64_t kr = kernel->rank >= 1 ? kernel->shape[0] : 1, kc = kernel->rank >= 2 ? kernel->shape[1] : 1;
  if (ni == 0 || nk == 0) {  // simple_provider.rs:6079-6093
    uint64_t es[2] = {mode == RM_CONV_SAME ? sr : 0, mode == RM_CONV_SAME ? sc : 0};
    return rm_zeros(p, es, 2, out);
  }
  // The host scatters out[r+i, c+j] += a[r,c] * ker[K-1-i, K-1-j] (conv2.rs:603-617 == simple_provider.rs:1856-1873). In
  // gather form that is out(o) = sum_u ker(u) * sig(o + start - (K-1) + u) with the kernel as stored, zero outside, then
  // the conv2 slicing (simple_provider.rs:1908-1956). The reference's KATs pin exactly this (conv2.rs:736-978). Visiting u
  // in column-major order equals the host's accumulation order (signal index ascending), so the result is bit-identical.
  FilterParams fp{};
  const uint64_t se[2] = {sr, sc}, ke[2] = {kr, kc};
  for (int d = 0; d < 3; ++d) { fp.ie[d] = d < 2 ? se[d] : 1; fp.ke[d] = d < 2 ? ke[d] : 1; fp.oe[d] = 1; fp.origin[d] = 0; fp.base[d] = 0; }
  for (int d = 0; d < 2; ++d) {
    int64_t start = 0;
    if (mode == RM_CONV_FULL) fp.oe[d] = se[d] + ke[d] - 1;
    else if (mode == RM_CONV_SAME) { fp.oe[d] = se[d]; start = (int64_t)((ke[d] - 1) / 2); }
    else { fp.oe[d] = se[d] >= ke[d] ? se[d] - ke[d] + 1 : 0; start = (int64_t)ke[d] - 1; }
    fp.base[d] = start - ((int64_t)ke[d] - 1);
  }
  if (mode == RM_CONV_VALID && (fp.oe[0] == 0 || fp.oe[1] == 0)) { uint64_t es[2] = {0, 0}; return rm_zeros(p, es, 2, out); }
  fp.padding = 0;
  fp.mode = 0;
  fp.cval = 0.0;
  uint64_t oshape[2] = {fp.oe[0], fp.oe[1]};
  void* po;
  RM_TRY(alloc_tensor(p, oshape, 2, out, &po));
  const uint64_t total = fp.oe[0] * fp.oe[1];
  if (total == 0) return RM_OK;
  launch_filter(p, pi, pk, po, total, fp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rm_free(p, out); return fail(RM_ERROR, "conv2d launch failed: %s", cudaGetErrorString(e)); }
  count_launch(p);
  return RM_OK;
}

static rm_status minmax_dim(rm_provider* p, const rm_handle* a, uint32_t dim, bool is_min, rm_handle* values, rm_handle* indices) {
  RM_REQUIRE(p && a && values && indices, RM_INVALID_ARG, "reduce_min/max_dim: bad arguments");
  const uint32_t rank = std::max<uint32_t>(a->rank, 2);
  RM_REQUIRE(dim < rank, RM_ERROR, "reduce_%s_dim: dim %u out of range", is_min ? "min" : "max", dim);
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, a, &src, nullptr));
  uint64_t pre = 1, n = 1, post = 1, oshape[RM_MAX_RANK];
  for (uint32_t d = 0; d < rank; ++d) {
    const uint64_t e = d < a->rank ? a->shape[d] : 1;
    if (d < dim) pre *= e; else if (d == dim) n = e; else post *= e;
    oshape[d] = d == dim ? 1 : e;
  }
  void *pv, *pidx;
  RM_TRY(alloc_tensor(p, oshape, rank, values, &pv));
  rm_status st = alloc_tensor(p, oshape, rank, indices, &pidx);
  if (st != RM_OK) { rm_free(p, values); return st; }
  const uint64_t slices = pre * post;
  if (slices == 0) return RM_OK;
  const uint64_t blocks = (slices * 32 + 255) / 256;
  if (p->precision == RM_F64) {
    if (is_min) minmax_dim_kernel<double, true><<<(unsigned)blocks, 256, 0, p->stream>>>((const double*)src, (double*)pv, (double*)pidx, pre, n, post);
    else minmax_dim_kernel<double, false><<<(unsigned)blocks, 256, 0, p->stream>>>((const double*)src, (double*)pv, (double*)pidx, pre, n, post);
  } else {
    if (is_min) minmax_dim_kernel<float, true><<<(unsigned)blocks, 256, 0, p->stream>>>((const float*)src, (float*)pv, (float*)pidx, pre, n, post);
    else minmax_dim_kernel<float, false><<<(unsigned)blocks, 256, 0, p->stream>>>((const float*)src, (float*)pv, (float*)pidx, pre, n, post);
  }
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}
RM_EXPORT rm_status rm_reduce_max_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* values, rm_handle* indices) { return minmax_dim(p, a, dim, false, values, indices); }
RM_EXPORT rm_status rm_reduce_min_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* values, rm_handle* indices) { return minmax_dim(p, a, dim, true, values, indices); }

RM_EXPORT rm_status rm_warmup(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  DeviceGuard g(p->ordinal);
  // pre-compile the most common fused programs (mirrors AccelProvider::warmup, lib.rs:3010)
  uint64_t shp[2] = {8, 8};
  rm_handle a, b, c;
  RM_TRY(rm_ones(p, shp, 2, &a));
  rm_status st = rm_elem_add(p, &a, &a, &b);
  if (st == RM_OK) { rm_free(p, &b); st = rm_elem_mul(p, &a, &a, &b); }
  if (st == RM_OK) { rm_free(p, &b); st = rm_reduce_sum(p, &a, &c); }
  if (st == RM_OK) rm_free(p, &c);
  rm_free(p, &a);
  if (st == RM_OK) { RM_CUDA(cudaStreamSynchronize(p->stream)); p->host_syncs.fetch_add(1, std::memory_order_relaxed); }
  return st;
}
