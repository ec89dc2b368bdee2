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

// C = lhs*rhs; every column divided by sqrt(sum(col.^2) + epsilon)   (simple_provider.rs:7852-7890)
// Three launches behind the GEMM-shaped pattern: the product, ONE generated reduction that squares while it sums the columns, ONE
// generated broadcast kernel that applies c / sqrt(ss + eps) (round 1 chained six single-operator kernels and four temporaries).
RM_EXPORT rm_status rm_matmul_power_step(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, double epsilon, rm_handle* out) {
  RM_REQUIRE(p && lhs && rhs && out, RM_INVALID_ARG, "matmul_power_step: bad arguments");
  DeviceGuard g(p->ordinal);
  ScopedWall wall(p->t_matmul);
  Temp c(p), ss(p), eps(p);
  RM_TRY(matmul_impl(p, lhs, rhs, nullptr, &c.h));
  c.live = true;
  const uint64_t rows = c.h.shape[0], cols = c.h.shape[1];
  if (rows * cols == 0) { *out = c.h; c.live = false; return RM_OK; }
  ReductionProgram sq;
  sq.scalar_ty = p->precision == RM_F64 ? "f64" : "f32";
  sq.n_inputs = 1;
  sq.axis = 0;
  sq.val_expr = "(v0 * v0)";
  uint64_t sshape[2] = {1, cols};
  RM_TRY(run_reduction_program(p, sq, "pattern:colsumsq", RedOp::Sum, RedLayout::Contig, &c.h, 1, sshape, 2, rows, cols, /*inner=*/cols, 0, 1.0, &ss.h));
  ss.live = true;
  uint64_t one[2] = {1, 1};
  RM_TRY(rm_fill(p, one, 2, epsilon, &eps.h));
  eps.live = true;
  ElementwiseProgram nrm;
  nrm.scalar_ty = sq.scalar_ty;
  nrm.n_inputs = 3;
  nrm.n_outputs = 1;
  nrm.outputs.push_back("(v0 / sqrt(v1 + v2))");  // norm = sqrt(sum + epsilon); value / norm
  rm_handle in[3] = {c.h, ss.h, eps.h};
  return run_elementwise_program(p, nrm, "pattern:power_step_normalize", in, 3, c.h.shape, 2, rows * cols, out);
}

// normalization: 0 = Unbiased (n-1), 1 = Biased (n). Rows = All, no weights (what the CenteredGram pattern requests).
// mean -> centre (one generated broadcast kernel) -> Xc' * Xc on the FP64 tensor-core GEMM reading Xc transposed in place, the
// division by (n - 1) fused into its store (round 1: materialised transpose + full GEMM + a separate scalar divide).
RM_EXPORT rm_status rm_covariance(rm_provider* p, const rm_handle* matrix, int normalization_biased, rm_handle* out) {
  RM_REQUIRE(p && matrix && out, RM_INVALID_ARG, "covariance: bad arguments");
  RM_REQUIRE(matrix->rank == 2, RM_ERROR, "covariance: matrix input required");
  DeviceGuard g(p->ordinal);
  const uint64_t rows = matrix->shape[0], cols = matrix->shape[1];
  uint64_t oshape[2] = {cols, cols};
  if (cols == 0) return rm_zeros(p, oshape, 2, out);
  const double denom = normalization_biased ? (double)rows : (double)rows - 1.0;
  if (!(denom > 0.0)) return rm_fill(p, oshape, 2, NAN, out);  // cov.rs:931-933
  Temp mean(p), xc(p), dv(p);
  RM_TRY(rm_reduce_mean_dim(p, matrix, 0, &mean.h));  // [1, cols]
  mean.live = true;
  RM_TRY(rm_elem_sub(p, matrix, &mean.h, &xc.h));     // centred, broadcast over rows
  xc.live = true;
  if (p->precision != RM_F64) {
    Temp gram(p);
    RM_TRY(syrk_impl(p, &xc.h, nullptr, &gram.h));
    gram.live = true;
    return rm_scalar_div(p, &gram.h, denom, out);
  }
  uint64_t vshape[2] = {1, cols};
  RM_TRY(rm_fill(p, vshape, 2, denom, &dv.h));
  dv.live = true;
  void* pdv;
  RM_TRY(resolve(p, &dv.h, &pdv, nullptr));
  return syrk_impl(p, &xc.h, (const double*)pdv, out);
}
