// patterns.cu — provider hooks behind the planner's special fusion kinds:
//   CenteredGram       -> covariance(matrix, None, None, {normalization, rows: All})   (fusion_exec.rs:630-673)
//   PowerStepNormalize -> matmul_power_step(lhs, rhs, {epsilon})                       (fusion_exec.rs:675-728)
//   ExplainedVariance  -> matmul / reshape / diag_extract                              (fusion_exec.rs:730-868)
// They are compositions of the provider's own kernels (tensor-core GEMM, fused reductions, broadcast elementwise), so the
// intermediates never leave the device. Host semantics: simple_provider.rs:7852-7890 (power step), :3281-3312 (diag_extract),
// runmat-runtime/src/builtins/stats/summary/cov.rs:916-960 (unweighted covariance).
#include "common.h"

using namespace rm;

namespace {
template <typename T>
__global__ void diag_extract_kernel(const T* __restrict__ m, uint64_t rows, uint64_t r0, uint64_t c0, uint64_t len, T* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) out[i] = m[(r0 + i) + (c0 + i) * rows];
}
// frees `h` and preserves the error message of a failed step
struct Temp {
  rm_provider* p;
  rm_handle h{};
  bool live = false;
  explicit Temp(rm_provider* p_) : p(p_) {}
  ~Temp() { if (live) { std::string m = last_error(); rm_free(p, &h); set_error("%s", m.c_str()); } }
};
}  // namespace

RM_EXPORT rm_status rm_diag_extract(rm_provider* p, const rm_handle* matrix, int64_t offset, rm_handle* out) {
  RM_REQUIRE(p && matrix && out, RM_INVALID_ARG, "diag: bad arguments");
  RM_REQUIRE(matrix->rank == 2, RM_ERROR, "diag: matrix input required");
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, matrix, &src, nullptr));
  const uint64_t rows = matrix->shape[0], cols = matrix->shape[1];
  const uint64_t r0 = offset < 0 ? (uint64_t)(-offset) : 0, c0 = offset > 0 ? (uint64_t)offset : 0;
  const uint64_t len = (r0 < rows && c0 < cols) ? std::min(rows - r0, cols - c0) : 0;
  uint64_t oshape[2] = {len, 1};
  void* dst;
  RM_TRY(alloc_tensor(p, oshape, 2, out, &dst));
  if (len == 0) return RM_OK;
  const unsigned grid = (unsigned)std::min<uint64_t>((len + 255) / 256, 1024);
  if (p->precision == RM_F64) diag_extract_kernel<double><<<grid, 256, 0, p->stream>>>((const double*)src, rows, r0, c0, len, (double*)dst);
  else diag_extract_kernel<float><<<grid, 256, 0, p->stream>>>((const float*)src, rows, r0, c0, len, (float*)dst);
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

// C = lhs*rhs; every column divided by sqrt(sum(col.^2) + epsilon)
RM_EXPORT rm_status rm_matmul_power_step(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, double epsilon, rm_handle* out) {
  RM_REQUIRE(p && lhs && rhs && out, RM_INVALID_ARG, "matmul_power_step: bad arguments");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_matmul);
  Temp c(p), sq(p), ss(p), se(p), nrm(p);
  RM_TRY(matmul_impl(p, lhs, rhs, nullptr, &c.h));
  c.live = true;
  RM_TRY(rm_elem_mul(p, &c.h, &c.h, &sq.h));
  sq.live = true;
  RM_TRY(rm_reduce_sum_dim(p, &sq.h, 0, &ss.h));  // [1, cols]
  ss.live = true;
  RM_TRY(rm_scalar_add(p, &ss.h, epsilon, &se.h));
  se.live = true;
  RM_TRY(rm_unary_sqrt(p, &se.h, &nrm.h));
  nrm.live = true;
  return rm_elem_div(p, &c.h, &nrm.h, out);  // broadcast [rows,cols] ./ [1,cols]
}

// normalization: 0 = Unbiased (n-1), 1 = Biased (n). Rows = All, no weights (what the CenteredGram pattern requests).
RM_EXPORT rm_status rm_covariance(rm_provider* p, const rm_handle* matrix, int normalization_biased, rm_handle* out) {
  RM_REQUIRE(p && matrix && out, RM_INVALID_ARG, "covariance: bad arguments");
  RM_REQUIRE(matrix->rank == 2, RM_ERROR, "covariance: matrix input required");
  DeviceGuard g(p->ordinal);
  const uint64_t rows = matrix->shape[0], cols = matrix->shape[1];
  uint64_t oshape[2] = {cols, cols};
  if (cols == 0) return rm_zeros(p, oshape, 2, out);
  const double denom = normalization_biased ? (double)rows : (double)rows - 1.0;
  if (!(denom > 0.0)) return rm_fill(p, oshape, 2, NAN, out);  // cov.rs:931-933
  Temp mean(p), xc(p), gram(p);
  RM_TRY(rm_reduce_mean_dim(p, matrix, 0, &mean.h));  // [1, cols]
  mean.live = true;
  RM_TRY(rm_elem_sub(p, matrix, &mean.h, &xc.h));     // centred, broadcast over rows
  xc.live = true;
  RM_TRY(rm_syrk(p, &xc.h, &gram.h));                 // Xc' * Xc on the tensor-core GEMM
  gram.live = true;
  return rm_scalar_div(p, &gram.h, denom, out);
}
