// debug_api.cpp — device-free introspection of the fused lowering (used by the CPU test-suite and by
// build() to pre-populate the on-disk cubin cache). Mirrors RUNMAT_DEBUG_DUMP_FUSED_WGSL (fusion_exec.rs:583).
#include <algorithm>
#include "common.h"

using namespace rm;

static rm_status copy_out(const std::string& s, char* buf, size_t buflen, size_t* needed) {
  if (needed) *needed = s.size() + 1;
  if (buf && buflen) {
    size_t n = std::min(buflen - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = '\0';
  }
  return RM_OK;
}

// variant: 0 flat, 1 broadcast
RM_EXPORT rm_status rm_debug_lower_elementwise(const char* shader, int variant, uint32_t scalar_mask, char* buf, size_t buflen, size_t* needed) {
  ElementwiseProgram prog;
  std::string err;
  if (!parse_elementwise_wgsl(shader, &prog, &err)) return fail(RM_COMPILE_ERROR, "%s", err.c_str());
  return copy_out(emit_elementwise_cuda(prog, variant ? EwVariant::Broadcast : EwVariant::Flat, scalar_mask), buf, buflen, needed);
}
// op: 0 sum, 1 prod, 2 max, 3 min; layout: -1 = from shader axis, 0 contiguous, 1 strided, 2 interleaved
RM_EXPORT rm_status rm_debug_lower_reduction(const char* shader, int op, int layout, char* buf, size_t buflen, size_t* needed, int* axis, int* omit_nan) {
  ReductionProgram prog;
  std::string err;
  if (!parse_reduction_wgsl(shader, &prog, &err)) return fail(RM_COMPILE_ERROR, "%s", err.c_str());
  if (axis) *axis = prog.axis;
  if (omit_nan) *omit_nan = prog.omit_nan ? 1 : 0;
  RedLayout l = layout < 0 ? (prog.axis == 0 ? RedLayout::Contig : RedLayout::Strided) : layout == 2 ? RedLayout::Interleaved : (layout ? RedLayout::Strided : RedLayout::Contig);
  return copy_out(emit_reduction_cuda(prog, (RedOp)op, l), buf, buflen, needed);
}
RM_EXPORT rm_status rm_debug_translate_expr(const char* wgsl_expr, const char* scalar_ty, char* buf, size_t buflen) {
  std::string out, err;
  if (!translate_expr(wgsl_expr, scalar_ty, &out, &err)) return fail(RM_COMPILE_ERROR, "%s", err.c_str());
  return copy_out(out, buf, buflen, nullptr);
}
// NVRTC-compiles `src` for sm_100a (no device required); returns the cubin size.
RM_EXPORT rm_status rm_debug_compile(const char* src, const char* name, size_t* cubin_size) {
  std::vector<char> cubin;
  std::string log;
  RM_TRY(compile_cuda_to_cubin(src, name ? name : "k.cu", &cubin, &log));
  if (cubin_size) *cubin_size = cubin.size();
  return RM_OK;
}
