// common.h — internal definitions shared by the provider's translation units.
// The public surface is include/rm_accel.h; nothing here is exported.
#pragma once
#include <utility>

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rm_accel.h"

#define RM_EXPORT extern "C" __attribute__((visibility("default")))

namespace rm {

// ---- error plumbing: every entry point returns rm_status; the message is thread-local (anyhow::Error) --
void set_error(const char* fmt, ...);
const char* last_error();

struct Status {
  rm_status code;
  Status(rm_status c = RM_OK) : code(c) {}
  bool ok() const { return code == RM_OK; }
};

rm_status fail(rm_status code, const char* fmt, ...);

#define RM_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      cudaGetLastError();                                                                          \
      return ::rm::fail(_e == cudaErrorMemoryAllocation ? RM_OOM : RM_ERROR, "%s failed: %s (%s:%d)", \
                        #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    }                                                                                              \
  } while (0)

#define RM_TRY(expr)                       \
  do {                                     \
    rm_status _s = (expr);                 \
    if (_s != RM_OK) return _s;            \
  } while (0)

#define RM_REQUIRE(cond, code, ...)                       \
  do {                                                    \
    if (!(cond)) return ::rm::fail((code), __VA_ARGS__);  \
  } while (0)

// ---- buffer table -----------------------------------------------------------------------------------------
struct Buffer {
  void* ptr = nullptr;
  uint64_t elems = 0;  // logical elements (real storage only)
  cudaEvent_t ready = nullptr;  // set by upload: recorded on the H2D stream; the compute stream waits for it at first use
  uint64_t p2p_step1 = 0;       // comm.cu: step + 1 of the peer-exchange combine that still has to be enqueued into this buffer (0 = none)
};

struct DispatchCounter {
  std::atomic<uint64_t> count{0};
  std::atomic<uint64_t> wall_ns{0};
  void record(uint64_t ns) { count.fetch_add(1, std::memory_order_relaxed); wall_ns.fetch_add(ns, std::memory_order_relaxed); }
  void reset() { count = 0; wall_ns = 0; }
};

struct ScopedWall {
  DispatchCounter& c;
  std::chrono::steady_clock::time_point t0;
  explicit ScopedWall(DispatchCounter& c_) : c(c_), t0(std::chrono::steady_clock::now()) {}
  ~ScopedWall() {
    c.record((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
  }
};

struct FusedCache;  // fused.cu

}  // namespace rm

// The opaque provider. One CUDA device, one stream, one buffer table (SURVEY.md §8b "Ownership").
struct rm_provider {
  int ordinal = 0;
  uint32_t device_id = 0;
  rm_precision precision = RM_F64;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  cudaDeviceProp prop{};
  cudaMemPool_t pool = nullptr;

  std::mutex mu;  // guards `buffers`
  std::unordered_map<uint64_t, rm::Buffer> buffers;
  std::atomic<uint64_t> next_id{1};
  std::atomic<uint64_t> live_bytes{0};

  // RNG (simple_provider.rs:62, :3627-3640): host-LCG-compatible stream
  std::mutex rng_mu;
  uint64_t rng_state = 0x9e3779b97f4a7c15ULL;

  // telemetry (accelerate/src/telemetry.rs:15-250)
  rm::DispatchCounter t_fused_elementwise, t_fused_reduction, t_matmul, t_linsolve, t_mldivide, t_mrdivide;
  std::atomic<uint64_t> upload_bytes{0}, download_bytes{0}, cache_hits{0}, cache_misses{0}, kernel_launches{0};

  // scratch. `scratch_mu` is held from ensure_scratch() through pointer capture to the launch that uses the scratch, so a
  // concurrent caller cannot re-allocate it underneath an enqueue (the stream then orders the kernels themselves).
  std::mutex scratch_mu;
  void* reduce_scratch = nullptr;   // partial sums + tickets for two-stage reductions
  size_t reduce_scratch_bytes = 0;
  // gemm_ozaki.cu workspace (slices, exponents, device flags, tensor maps); oz_mu is held across one product's enqueues
  std::mutex oz_mu;
  void* oz_ws = nullptr;
  // number of times an entry point blocked the host on the device (download, read_scalar, synchronize, find, mldivide's gate ...)
  std::atomic<uint64_t> host_syncs{0};
  void* dev_flags = nullptr;  // int[64], zero-initialised: [0] imfilter TMA pipeline protocol error (image.cu)
  void* l2_flush = nullptr;
  cudaStream_t aux_stream = nullptr;  // solve.cu: update stream of the look-ahead LU (created at first use)
  size_t l2_flush_bytes = 0;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  // uploads run on their own stream so H2D copies overlap compute and D2H traffic queued on `stream`
  cudaStream_t h2d_stream = nullptr;
  std::mutex ev_mu;
  std::vector<cudaEvent_t> event_pool;  // recycled per-buffer "ready" events
  int matmul_engine = 0;
  bool launch_overlap = true;  // fused.cu: programmatic dependent launch of the generated kernels (rm_set_launch_overlap)
  // multi-GPU exchange (comm.cu): one NCCL communicator per provider, collectives on their own stream
  void* nccl_comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  int comm_rank = 0, comm_world = 1;
  std::mutex comm_mu;   // one exchange (publish + combine enqueue) at a time
  void* p2p = nullptr;  // comm.cu: peer-memory slot exchange state

  // bounded launch log (ProviderTelemetry::kernel_launches): ring of the last RM_MAX_KERNEL_LAUNCH_EVENTS events
  std::mutex log_mu;
  rm_kernel_launch_event launch_log[RM_MAX_KERNEL_LAUNCH_EVENTS];
  uint64_t launch_log_n = 0;  // events recorded so far (ring index = n % cap)

  rm::FusedCache* fused = nullptr;

  size_t elem_size() const { return precision == RM_F64 ? 8 : 4; }
};

namespace rm {

// RAII device selection: every entry point runs with the provider's device current.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int ordinal) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != ordinal) { cudaSetDevice(ordinal); changed = true; }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

inline uint64_t shape_elems(const uint64_t* shape, uint32_t rank) {
  uint64_t n = 1;
  for (uint32_t i = 0; i < rank; ++i) n *= shape[i];
  return n;
}
inline uint64_t handle_elems(const rm_handle* h) { return shape_elems(h->shape, h->rank); }

// Allocates a device buffer of `elems` elements (provider precision) and fills `out` with a fresh handle.
rm_status alloc_tensor(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out, void** dptr);
void comm_destroy(rm_provider* p);
// comm.cu peer-memory exchange: parameters of the publish tail fused into a producing kernel
struct P2PPublish {
  void* const* peers = nullptr;  // device array of per-rank slot buffers
  uint32_t n = 0, rank = 0;
  uint64_t step = 0;
  // combine of an earlier step fused in front of this publish (bank-reuse rule of the protocol): destination of the folded sum
  // (provider precision), its step + 1 (0 = nothing pending) and the device error flag of the bounded wait
  void* prev_dst = nullptr;
  uint64_t prev_step1 = 0;
  int* err = nullptr;
};
bool p2p_begin(rm_provider* p, P2PPublish* pub);     // caller holds p->comm_mu until p2p_finish()
rm_status p2p_finish(rm_provider* p, rm_handle* out);
// Enqueues the combine of exchange step `step` into `dst` on the compute stream. Caller holds p->mu (the buffer-table lock), so the
// combine is ordered before any consumer another host thread may enqueue for the same buffer.
void p2p_enqueue_combine_locked(rm_provider* p, uint64_t step, void* dst);
// Resolves a handle to its device pointer, validating device_id and element count.
rm_status resolve(rm_provider* p, const rm_handle* h, void** dptr, uint64_t* elems);
rm_status ensure_scratch(rm_provider* p, size_t bytes);
#if defined(__CUDACC__)
// Programmatic dependent launch for the hand-written kernels (the generated ones get it in fused.cu). A kernel launched through
// launch_pdl() opens with pdl_prologue(): it lets the NEXT kernel of the stream become resident as soon as SM resources allow and
// then waits until the PREVIOUS kernel has completed and its writes are visible -- before it touches global memory, so stream order
// is kept exactly; what overlaps is launch latency, block scheduling and the smem/barrier set-up of one kernel with the drain of
// the one before. Both instructions are no-ops in a plain launch; rm_set_launch_overlap(0) turns the attribute off.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool overlap, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif
inline void count_launch(rm_provider* p, uint64_t n = 1) { p->kernel_launches.fetch_add(n, std::memory_order_relaxed); }
struct LaunchAttr { const char* key; uint64_t value; };
// record_kernel_launch (accelerate/src/telemetry.rs:219-239): newest last, oldest dropped beyond the cap
inline void record_launch(rm_provider* p, const char* kernel, std::initializer_list<LaunchAttr> shape, std::initializer_list<LaunchAttr> tuning) {
  std::lock_guard<std::mutex> lk(p->log_mu);
  rm_kernel_launch_event& e = p->launch_log[p->launch_log_n % RM_MAX_KERNEL_LAUNCH_EVENTS];
  memset(&e, 0, sizeof e);
  snprintf(e.kernel, sizeof e.kernel, "%s", kernel);
  e.precision = p->precision;
  for (const LaunchAttr& a : shape) if (e.n_shape < RM_LAUNCH_ATTRS) { snprintf(e.shape[e.n_shape].key, 16, "%s", a.key); e.shape[e.n_shape++].value = a.value; }
  for (const LaunchAttr& a : tuning) if (e.n_tuning < RM_LAUNCH_ATTRS) { snprintf(e.tuning[e.n_tuning].key, 16, "%s", a.key); e.tuning[e.n_tuning++].value = a.value; }
  ++p->launch_log_n;
}

// normalize_scalar_shape-style helper: MATLAB values are rank >= 2.
inline void make_shape2(uint64_t r, uint64_t c, uint64_t* shape) { shape[0] = r; shape[1] = c; }

#define RM_LAUNCH_CHECK()                                                                          \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) return ::rm::fail(RM_ERROR, "kernel launch failed: %s (%s:%d)",         \
                                             cudaGetErrorString(_e), __FILE__, __LINE__);          \
  } while (0)

// ---- fused program lowering (fusion_lower.cpp) ---------------------------------------------------------
struct ElementwiseProgram {
  std::string scalar_ty;             // "f64" | "f32"
  uint32_t n_inputs = 0;
  uint32_t n_outputs = 0;
  std::vector<std::string> stmts;    // "T tmpK = <cuda expr>;"
  std::vector<std::string> outputs;  // cuda expr per output
};
struct ReductionProgram {
  std::string scalar_ty;
  uint32_t n_inputs = 0;
  int axis = 0;        // 0: reduce over rows of a column-major [nrows x ncols]; 1: reduce over cols
  bool omit_nan = false;
  std::string val_expr;  // cuda expr over v, v1, v2...
};
// Parse the reference planner's WGSL (fusion.rs:1525-1763 / :1765-2077). Returns false + message on failure.
bool parse_elementwise_wgsl(const char* shader, ElementwiseProgram* prog, std::string* err);
bool parse_reduction_wgsl(const char* shader, ReductionProgram* prog, std::string* err);
// Translate one WGSL expression to CUDA C (identifier/function rewrite, literal typing).
bool translate_expr(const std::string& wgsl, const std::string& scalar_ty, std::string* cuda, std::string* err);

enum class EwVariant { Flat = 0, Broadcast = 1 };
// scalar_mask bit k set => input k is a 1-element tensor (hoisted scalar load) in the Flat variant.
std::string emit_elementwise_cuda(const ElementwiseProgram& prog, EwVariant variant, uint32_t scalar_mask);
enum class RedOp { Sum = 0, Prod = 1, Max = 2, Min = 3 };
// Contig: slice s at [s*len, (s+1)*len); Strided: elem r of slice s at s + r*num_slices; Interleaved: the Strided addressing
// for a small power-of-two `inner`, swept as ONE flat 256-bit vector stream per [inner x len] block (vector lane l of a
// thread always belongs to the same slice, so the per-slice accumulators live in registers).
enum class RedLayout { Contig = 0, Strided = 1, Interleaved = 2 };
std::string emit_reduction_cuda(const ReductionProgram& prog, RedOp op, RedLayout layout);

// ---- fused module cache / launch (fused.cu) -----------------------------------------------------------------
rm_status fused_cache_create(rm_provider* p);
void fused_cache_destroy(rm_provider* p);
rm_status run_elementwise_program(rm_provider* p, const ElementwiseProgram& prog, const std::string& key,
                                  const rm_handle* inputs, uint32_t n_inputs, const uint64_t* out_shape,
                                  uint32_t rank, uint64_t len, rm_handle* outs);
// Runs a reduction program. `out_shape/rank` describe the result handle. scale: result = use_div ? acc/factor : acc*factor.
// Strided layout: element r of slice s lives at (s % inner) + (s / inner) * inner * reduce_len + r * inner.
rm_status run_reduction_program(rm_provider* p, const ReductionProgram& prog, const std::string& key, RedOp op,
                                RedLayout layout, const rm_handle* inputs, uint32_t n_inputs,
                                const uint64_t* out_shape, uint32_t rank, uint64_t reduce_len,
                                uint64_t num_slices, uint64_t inner, int use_div, double factor, rm_handle* out,
                                const struct P2PPublish* publish = nullptr, double param0 = 0.0);
// NVRTC-only compile (no device needed): used by CPU tests of the lowering and by rm_warmup.
rm_status compile_cuda_to_cubin(const std::string& src, const char* name, std::vector<char>* cubin, std::string* log);

// ---- other translation units ---------------------------------------------------------------------------------
rm_status matmul_impl(rm_provider* p, const rm_handle* a, const rm_handle* b, const rm_matmul_epilogue* ep, rm_handle* out);
// out = a' * a, every entry divided by divisor_vec[col] when given (device vector of a->shape[1] doubles); no materialised transpose on f64
rm_status syrk_impl(rm_provider* p, const rm_handle* a, const double* divisor_vec, rm_handle* out);
// gemm_ozaki.cu: f64 GEMM on tcgen05 (int8 Ozaki split). *used=false => shape outside the engine's range, caller runs the
// DMMA engine unconditionally; *used=true => caller enqueues the DMMA kernel CONDITIONALLY on `guard` (device flags: non-finite
// inputs or tiles that failed the element-wise accuracy guard), while still holding p->oz_mu.
struct OzGuard {
  const int* flags = nullptr;      // flags[0] != 0: recompute every tile
  const int* tileflags = nullptr;  // per 128x256 Ozaki tile, m fastest
  int tiles_m = 0;
};
rm_status ozaki_matmul(rm_provider* p, const double* A, const double* B, double* C, uint64_t m, uint64_t n, uint64_t k,
                       const rm_matmul_epilogue* epd, const void* prow, const void* pcol, void* pdiag, bool ep_active, bool* used, OzGuard* guard);
int ozaki_default_slices();
void ozaki_workspace_destroy(rm_provider* p);
rm_status ozaki_last_stats(rm_provider* p, int out[4]);
rm_status dgemm_sub_strided(rm_provider* p, const double* A, uint64_t lda, const double* B, uint64_t ldb, double* C, uint64_t ldc,
                            uint64_t m, uint64_t n, uint64_t k, cudaStream_t stream = nullptr);

}  // namespace rm
