/*
 * rm_accel.h — C ABI of the B200-native accelerate provider for RunMat's dense-array hot path.
 *
 * Every entry point is the C image of one method of the reference trait
 *   `runmat_accelerate_api::AccelProvider`  (crates/runmat-accelerate-api/src/lib.rs:1386-3152)
 * restricted to the hot path (SURVEY.md §8 rows a1–a15). A Rust shim (`shim/cuda_provider.rs`)
 * forwards `impl AccelProvider for CudaProvider` to these functions 1:1; see INTEGRATION.md.
 *
 * Conventions (mirroring the trait's contract):
 *  - Tensors are dense, column-major. Host data crosses the boundary as `const double*` (upload) and
 *    `double*` (download): the host `Tensor` always stores f64 (runmat-builtins/src/lib.rs:425-436).
 *    Device storage is f64 (provider precision F64, the default) or f32 (precision F32: upload narrows,
 *    download widens — the same rule as the wgpu provider, backend/wgpu/provider/ops/io.rs:87).
 *  - A handle is the POD image of `GpuTensorHandle{shape,device_id,buffer_id}` (lib.rs:260-264). The
 *    handle's shape is authoritative (the trait's default `reshape` only edits it, lib.rs:2676-2684);
 *    the provider checks prod(shape) against the buffer's element count.
 *  - Every call returns an `rm_status`. Non-zero mirrors `Err(anyhow!(..))`; the message is available
 *    from `rm_last_error()` (thread-local). `RM_UNSUPPORTED` mirrors the trait's default
 *    "... not supported by provider" so the caller can fall back to host exactly as it does today.
 *    No call aborts the process; device OOM is reported as `RM_OOM`.
 *  - Results are NEW handles, except `rm_scatter_linear` (in place, lib.rs:1438) and the
 *    `diag_output` of `rm_matmul_epilogue` (written in place, lib.rs:3520-3522).
 *  - The provider is `Send + Sync`: all entry points are thread-safe (one mutex around the buffer
 *    table, stream-ordered allocation); `rm_free` may arrive from a GC finalizer thread.
 *  - Work is enqueued on the provider's CUDA stream. The calls that wait for the device are the ones whose
 *    return value lives on it: `rm_download*`, `rm_read_scalar`, `rm_synchronize`, `rm_timer_end_ms`,
 *    `rm_warmup`, `rm_find` (data-dependent output size), `rm_sub2ind`/`rm_ind2sub` (the host raises on a bad
 *    subscript) and `rm_mldivide`/`rm_mrdivide`/`rm_linsolve` (the conditioning gate decides between a result
 *    and `RM_UNSUPPORTED`). Everything else — including `rm_matmul` on either engine — only enqueues;
 *    `rm_host_sync_count` counts the waits so a caller (and the tests) can verify that.
 *  - There is NO CPU fallback inside the library: without a usable CUDA device
 *    `rm_provider_create` fails with `RM_NO_DEVICE`.
 */
#ifndef RM_ACCEL_H
#define RM_ACCEL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RM_MAX_RANK 16
#define RM_ABI_VERSION 1

typedef enum rm_status {
  RM_OK = 0,
  RM_ERROR = 1,          /* generic anyhow::Error */
  RM_UNSUPPORTED = 2,    /* "... not supported by provider" */
  RM_OOM = 3,
  RM_INVALID_HANDLE = 4, /* unknown buffer / foreign device_id (simple_provider.rs:2753-2760) */
  RM_INVALID_ARG = 5,
  RM_NO_DEVICE = 6,
  RM_COMPILE_ERROR = 7   /* fused program could not be lowered/compiled */
} rm_status;

/* ProviderPrecision (lib.rs:815-819) */
typedef enum rm_precision { RM_F32 = 0, RM_F64 = 1 } rm_precision;

/* GpuTensorHandle (lib.rs:260-264) */
typedef struct rm_handle {
  uint64_t buffer_id;
  uint32_t device_id;
  uint32_t rank;
  uint64_t shape[RM_MAX_RANK];
} rm_handle;

typedef struct rm_provider rm_provider; /* opaque: one CUDA device + stream + buffer table */

/* ApiDeviceInfo (lib.rs:1448-1456) */
typedef struct rm_device_info {
  uint32_t device_id;
  char name[128];
  char vendor[32];
  char backend[32];
  uint64_t memory_bytes;
  uint32_t sm_count;
  uint32_t cc_major, cc_minor;
} rm_device_info;

/* ProviderDispatchStats / ProviderTelemetry (lib.rs:1337-1376) */
typedef struct rm_dispatch_stats {
  uint64_t count;
  uint64_t total_wall_time_ns;
} rm_dispatch_stats;

typedef struct rm_telemetry {
  rm_dispatch_stats fused_elementwise;
  rm_dispatch_stats fused_reduction;
  rm_dispatch_stats matmul;
  rm_dispatch_stats linsolve;
  rm_dispatch_stats mldivide;
  rm_dispatch_stats mrdivide;
  uint64_t upload_bytes;
  uint64_t download_bytes;
  uint64_t fusion_cache_hits;
  uint64_t fusion_cache_misses;
  uint64_t kernel_launches; /* total device kernels enqueued by this provider */
} rm_telemetry;

/* ---- lifecycle / registry (register_provider, lib.rs:3213-3273) ---------------------------------- */
/* cuda_ordinal: CUDA device index; device_id: the id `next_device_id()` handed out (lib.rs:3279). */
rm_status rm_provider_create(int cuda_ordinal, uint32_t device_id, rm_precision precision,
                             rm_provider** out);
rm_status rm_provider_destroy(rm_provider* p);
uint32_t rm_abi_version(void);
const char* rm_last_error(void);
rm_status rm_device_info_string(rm_provider* p, char* buf, size_t buflen); /* device_info() :1390 */
rm_status rm_device_info_struct(rm_provider* p, rm_device_info* out);      /* :1448 */
/* ---- multi-GPU exchange (no counterpart in the single-device trait) --------------------------------------
 * One process per GPU. Rank 0 calls rm_comm_unique_id, the host ships the 128 bytes to every rank (torch.distributed,
 * MPI, a file ...), every rank calls rm_comm_init. rm_comm_allreduce_sum(in) -> out is stream-ordered after the producer of
 * `in`, runs on a communication stream, and `out` is waited for lazily at its first use. NCCL is dlopen'ed on first use. */
#define RM_COMM_ID_BYTES 128
rm_status rm_comm_unique_id(uint8_t* out, uint32_t len);
rm_status rm_comm_init(rm_provider* p, const uint8_t* unique_id, uint32_t len, uint32_t rank, uint32_t world);
uint32_t rm_comm_world_size(rm_provider* p);
rm_status rm_comm_allreduce_sum(rm_provider* p, const rm_handle* in, rm_handle* out);
rm_status rm_comm_fence(rm_provider* p);
/* Peer-memory exchange for the scalar of a sharded reduction (NVLink/NVSwitch peer stores; comm.cu). Every rank calls
 * rm_comm_p2p_export (allocates its slot buffer, returns the 64-byte CUDA IPC handle), the host gathers the handles of all
 * ranks in rank order, every rank calls rm_comm_p2p_connect. Afterwards rm_comm_allreduce_sum of a 1-element f64 tensor and
 * rm_fused_reduction_allreduce use peer stores instead of a NCCL launch; results are bit-identical on every rank (fixed rank
 * order). The fold ("combine") of an exchange is lazy: it is fused in front of the publish issued four exchanges later, or
 * enqueued on the provider stream when the result handle is first used or freed -- no communication stream, no kernel of its
 * own in a steady loop. Every rank must issue the same sequence of exchanges. world == 1 connects a rank to itself. */
#define RM_COMM_P2P_HANDLE_BYTES 64
rm_status rm_comm_p2p_export(rm_provider* p, uint8_t* handle_out, uint32_t len);
rm_status rm_comm_p2p_connect(rm_provider* p, const uint8_t* all_handles, uint32_t len, uint32_t rank, uint32_t world);
int rm_comm_p2p_connected(rm_provider* p);
rm_status rm_comm_p2p_error(rm_provider* p, int32_t* err); /* 1 if a combine's bounded wait expired (a peer never published) */
/* PCI bus id ("0000:1b:00.0") of the provider's device: lets the host place its threads and pinned buffers on the GPU's NUMA node. */
rm_status rm_device_pci_bus_id(rm_provider* p, char* buf, uint32_t buflen);
uint32_t rm_device_id(rm_provider* p);                                     /* :1391 */
rm_precision rm_provider_precision(rm_provider* p);                        /* precision() :1458 */
rm_status rm_synchronize(rm_provider* p);
uint64_t rm_host_sync_count(rm_provider* p); /* number of entry-point calls that blocked the host on the device so far */
/* export_context (lib.rs:1406): share the CUDA stream / raw device pointer with other CUDA code. */
rm_status rm_get_stream(rm_provider* p, void** cuda_stream_out);
rm_status rm_set_stream(rm_provider* p, void* cuda_stream);
rm_status rm_device_ptr(rm_provider* p, const rm_handle* h, void** dptr_out, uint64_t* elems_out);
rm_status rm_warmup(rm_provider* p);                                       /* warmup() :3010 */
/* stream-ordered D2D copy of a tensor into caller-owned device memory (feeds the final NCCL all-reduce without a host hop) */
rm_status rm_copy_to_device(rm_provider* p, const rm_handle* h, void* dst_device_ptr, uint64_t dst_elems);

/* ---- a2: upload / download / free (lib.rs:1387-1389) -------------------------------------------- */
rm_status rm_upload(rm_provider* p, const double* data, const uint64_t* shape, uint32_t rank,
                    rm_handle* out);
/* extension for hosts that already hold f32 (the 4K-image config): no f64 staging copy */
rm_status rm_upload_f32(rm_provider* p, const float* data, const uint64_t* shape, uint32_t rank,
                        rm_handle* out);
rm_status rm_download(rm_provider* p, const rm_handle* h, double* out, uint64_t out_len);
rm_status rm_download_f32(rm_provider* p, const rm_handle* h, float* out, uint64_t out_len);
/* stream-ordered D2H without the final wait: `out` is valid after rm_synchronize() (pipelined host<->device steps) */
rm_status rm_download_async(rm_provider* p, const rm_handle* h, double* out, uint64_t out_len);
rm_status rm_free(rm_provider* p, const rm_handle* h);
rm_status rm_read_scalar(rm_provider* p, const rm_handle* h, uint64_t linear_index, double* out); /* :1463 */
uint64_t rm_live_buffers(rm_provider* p);
uint64_t rm_live_bytes(rm_provider* p);

/* ---- a14: constructors, reshape, layout, indexing ------------------------------------------------ */
rm_status rm_zeros(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out);   /* :1468 */
rm_status rm_ones(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out);    /* :1492 */
rm_status rm_fill(rm_provider* p, const uint64_t* shape, uint32_t rank, double value, rm_handle* out); /* :1502 */
rm_status rm_eye(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out);     /* :1552 */
rm_status rm_linspace(rm_provider* p, double start, double stop, uint64_t count, rm_handle* out); /* :1887 */
rm_status rm_reshape(rm_provider* p, const rm_handle* h, const uint64_t* new_shape, uint32_t rank,
                     rm_handle* out);                                                       /* :2676 */
rm_status rm_transpose(rm_provider* p, const rm_handle* a, rm_handle* out);                 /* :2532 */
rm_status rm_permute(rm_provider* p, const rm_handle* a, const uint32_t* order, uint32_t n_order,
                     rm_handle* out);                                                       /* :2579 */
/* indices are zero-based u32 logical element indices (lib.rs:1423-1445) */
rm_status rm_gather_linear(rm_provider* p, const rm_handle* source, const uint32_t* indices,
                           uint64_t n_indices, const uint64_t* out_shape, uint32_t out_rank,
                           rm_handle* out);
rm_status rm_scatter_linear(rm_provider* p, const rm_handle* target, const uint32_t* indices,
                            uint64_t n_indices, const rm_handle* values);
rm_status rm_repmat(rm_provider* p, const rm_handle* a, const uint64_t* reps, uint32_t n_reps,
                    rm_handle* out);                                                        /* :2689 */
/* cat (lib.rs:2686): concatenate along the 1-based dimension `dim` */
rm_status rm_cat(rm_provider* p, uint32_t dim_one_based, const rm_handle* inputs, uint32_t n_inputs, rm_handle* out);

/* ---- indexing class of the "next" rows; all results are exact index arithmetic ------------ */
/* find (lib.rs:2937): 1-based linear/row/col indices + values of the non-zeros, ascending (direction_last=0) or the last
 * ones in descending order (direction_last=1); has_limit/limit = Option<usize>. Outputs are [count,1]. */
rm_status rm_find(rm_provider* p, const rm_handle* a, int has_limit, uint64_t limit, int direction_last,
                  rm_handle* linear, rm_handle* rows, rm_handle* cols, rm_handle* values);
rm_status rm_scatter_column(rm_provider* p, const rm_handle* matrix, uint64_t col_index, const rm_handle* values, rm_handle* out); /* :3064 */
rm_status rm_scatter_row(rm_provider* p, const rm_handle* matrix, uint64_t row_index, const rm_handle* values, rm_handle* out);    /* :3075 */
/* sub2ind (lib.rs:3084): subscripts are 1-based doubles; scalar_mask[d] != 0 broadcasts input d; errors like the host's
 * coerce_sub2ind_value (simple_provider.rs:2268-2291) */
rm_status rm_sub2ind(rm_provider* p, const uint64_t* dims, const uint64_t* strides, uint32_t ndims, const rm_handle* inputs,
                     const uint8_t* scalar_mask, uint64_t len, const uint64_t* output_shape, uint32_t rank, rm_handle* out);
rm_status rm_ind2sub(rm_provider* p, const uint64_t* dims, const uint64_t* strides, uint32_t ndims, const rm_handle* indices,
                     uint64_t total, uint64_t len, const uint64_t* output_shape, uint32_t rank, rm_handle* outs); /* :3102; outs[ndims] */

/* ---- a5: unfused operator surface (lib.rs:1890-2357) --------------------------------------------- */
typedef enum rm_binary_op {
  RM_BIN_ADD = 0, RM_BIN_SUB, RM_BIN_MUL, RM_BIN_DIV, RM_BIN_POW, RM_BIN_MAX, RM_BIN_MIN,
  RM_BIN_HYPOT, RM_BIN_ATAN2, RM_BIN_MOD, RM_BIN_REM,
  RM_BIN_GE, RM_BIN_LE, RM_BIN_LT, RM_BIN_GT, RM_BIN_EQ, RM_BIN_NE,
  RM_BIN__COUNT
} rm_binary_op;

typedef enum rm_unary_op {
  RM_UN_SIN = 0, RM_UN_COS, RM_UN_TAN, RM_UN_ASIN, RM_UN_ACOS, RM_UN_ATAN,
  RM_UN_SINH, RM_UN_COSH, RM_UN_TANH, RM_UN_ASINH, RM_UN_ACOSH, RM_UN_ATANH,
  RM_UN_EXP, RM_UN_EXPM1, RM_UN_LOG, RM_UN_LOG2, RM_UN_LOG10, RM_UN_LOG1P, RM_UN_SQRT,
  RM_UN_ABS, RM_UN_SIGN, RM_UN_FLOOR, RM_UN_CEIL, RM_UN_ROUND, RM_UN_FIX, RM_UN_NEG,
  RM_UN_POW2, RM_UN_HEAVISIDE, RM_UN_SINGLE, RM_UN_DOUBLE,
  RM_UN_ISNAN, RM_UN_ISINF, RM_UN_ISFINITE, RM_UN_NAN_TO_ZERO, RM_UN_NOT_NAN_MASK,
  RM_UN_ERF, RM_UN_GAMMA, RM_UN_GAMMALN,   /* unary_erf :2101, unary_gamma :2089, unary_gammaln :2095 */
  RM_UN__COUNT
} rm_unary_op;

typedef enum rm_scalar_op {
  RM_SC_ADD = 0, RM_SC_SUB, RM_SC_MUL, RM_SC_DIV, RM_SC_RSUB, RM_SC_RDIV, RM_SC_MAX, RM_SC_MIN,
  RM_SC_POW, RM_SC__COUNT
} rm_scalar_op;

/* generic dispatchers (MATLAB implicit expansion, builtins/common/broadcast.rs:8-176) */
rm_status rm_elem_binary(rm_provider* p, rm_binary_op op, const rm_handle* a, const rm_handle* b, rm_handle* out);
rm_status rm_unary(rm_provider* p, rm_unary_op op, const rm_handle* a, rm_handle* out);
rm_status rm_scalar_op_apply(rm_provider* p, rm_scalar_op op, const rm_handle* a, double scalar, rm_handle* out);

/* named trait images */
rm_status rm_elem_add(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1890 */
rm_status rm_elem_mul(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1897 */
rm_status rm_elem_max(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1904 */
rm_status rm_elem_min(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1911 */
rm_status rm_elem_sub(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1918 */
rm_status rm_elem_div(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1925 */
rm_status rm_elem_pow(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);   /* :1932 */
rm_status rm_elem_hypot(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out); /* :1979 */
rm_status rm_elem_atan2(rm_provider* p, const rm_handle* y, const rm_handle* x, rm_handle* out); /* :2069 */
rm_status rm_unary_sin(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2077 */
rm_status rm_unary_cos(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2211 */
rm_status rm_unary_tan(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2155 */
rm_status rm_unary_tanh(rm_provider* p, const rm_handle* a, rm_handle* out);   /* :2161 */
rm_status rm_unary_exp(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2259 */
rm_status rm_unary_log(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2271 */
rm_status rm_unary_sqrt(rm_provider* p, const rm_handle* a, rm_handle* out);   /* :2295 */
rm_status rm_unary_abs(rm_provider* p, const rm_handle* a, rm_handle* out);    /* :2241 */
rm_status rm_unary_floor(rm_provider* p, const rm_handle* a, rm_handle* out);  /* :2185 */
rm_status rm_unary_round(rm_provider* p, const rm_handle* a, rm_handle* out);  /* :2191 */
rm_status rm_scalar_add(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2340 */
rm_status rm_scalar_sub(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2343 */
rm_status rm_scalar_mul(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2346 */
rm_status rm_scalar_div(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2355 */
rm_status rm_scalar_rsub(rm_provider* p, const rm_handle* a, double s, rm_handle* out); /* :2333 */
rm_status rm_scalar_rdiv(rm_provider* p, const rm_handle* a, double s, rm_handle* out); /* :2336 */
rm_status rm_scalar_max(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2349 */
rm_status rm_scalar_min(rm_provider* p, const rm_handle* a, double s, rm_handle* out);  /* :2352 */

/* ---- a3: fused elementwise (lib.rs:2946-2977) ---------------------------------------------------- */
/* `shader` is the WGSL text the reference planner emits (fusion.rs:1525-1763). It is NOT compiled
 * as WGSL: the provider parses the closed `let tmpN: T = <expr>;` / `output.data[g] = <expr>;`
 * grammar (the same subset the reference's TestProvider re-parses, runmat-vm/tests/fusion_gpu.rs:886-949),
 * lowers it to a CUDA C kernel for sm_100a with -fmad=false, and caches the module by text hash. */
rm_status rm_fused_elementwise(rm_provider* p, const char* shader, const rm_handle* inputs,
                               uint32_t n_inputs, const uint64_t* output_shape, uint32_t rank,
                               uint64_t len, rm_handle* out);
rm_status rm_fused_elementwise_multi(rm_provider* p, const char* shader, const rm_handle* inputs,
                                     uint32_t n_inputs, const uint64_t* output_shape, uint32_t rank,
                                     uint64_t len, uint32_t num_outputs, rm_handle* outs);

/* ---- a4: fused reduction (lib.rs:2996-3007), ReductionFlavor (lib.rs:865-888) --------------------- */
typedef enum rm_reduction_flavor { RM_FLAVOR_SUM = 0, RM_FLAVOR_MEAN = 1, RM_FLAVOR_CUSTOM = 2 } rm_reduction_flavor;
rm_status rm_fused_reduction(rm_provider* p, const char* shader, const rm_handle* inputs,
                             uint32_t n_inputs, const uint64_t* output_shape, uint32_t rank,
                             uint64_t reduce_len, uint64_t num_slices, uint32_t workgroup_size,
                             rm_reduction_flavor flavor, double custom_scale, rm_handle* out);
void rm_fused_cache_counters(rm_provider* p, uint64_t* hits, uint64_t* misses); /* :3013 */
/* sharded 'all' reduction: per-rank fused reduction to one scalar + sum over the ranks of the peer-memory exchange;
 * the publish is the tail of the reduction kernel's last block. `out` is 1x1 and holds the GLOBAL value. */
rm_status rm_fused_reduction_allreduce(rm_provider* p, const char* shader, const rm_handle* inputs, uint32_t n_inputs,
                                       uint64_t reduce_len, rm_reduction_flavor flavor, double custom_scale, rm_handle* out);

/* ---- a6: reductions (lib.rs:2709-2882) ----------------------------------------------------------- */
typedef enum rm_nan_mode { RM_NAN_INCLUDE = 0, RM_NAN_OMIT = 1 } rm_nan_mode;
rm_status rm_reduce_sum(rm_provider* p, const rm_handle* a, rm_handle* out);                  /* :2709 */
rm_status rm_reduce_sum_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* out);/* :2715, dim 0-based */
rm_status rm_reduce_prod(rm_provider* p, const rm_handle* a, rm_handle* out);                 /* :2743 */
rm_status rm_reduce_mean(rm_provider* p, const rm_handle* a, rm_handle* out);                 /* :2756 */
rm_status rm_reduce_mean_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* out);/* :2779 */
rm_status rm_reduce_mean_nd(rm_provider* p, const rm_handle* a, const uint32_t* dims_zero_based,
                            uint32_t n_dims, rm_handle* out);                                 /* :2763 */
rm_status rm_reduce_moments_nd(rm_provider* p, const rm_handle* a, const uint32_t* dims_zero_based,
                               uint32_t n_dims, rm_handle* mean_out, rm_handle* ex2_out);     /* :2772 */
rm_status rm_reduce_max(rm_provider* p, const rm_handle* a, rm_handle* out);                  /* :2871 */
rm_status rm_reduce_min(rm_provider* p, const rm_handle* a, rm_handle* out);                  /* :2858 */
/* ReduceDimResult{values,indices}: indices are 1-based doubles (MATLAB) */
rm_status rm_reduce_max_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* values, rm_handle* indices); /* :2877 */
rm_status rm_reduce_min_dim(rm_provider* p, const rm_handle* a, uint32_t dim, rm_handle* values, rm_handle* indices); /* :2864 */
uint32_t rm_default_reduction_workgroup_size(rm_provider* p);                                 /* :3048 */
uint64_t rm_two_pass_threshold(rm_provider* p);                                               /* :3053 */

/* ---- a7/a8: matmul (+epilogue) (lib.rs:2375-2406, 3498-3550) -------------------------------------- */
typedef enum rm_scale_op { RM_SCALE_MULTIPLY = 0, RM_SCALE_DIVIDE = 1 } rm_scale_op;
typedef struct rm_matmul_epilogue {
  double alpha, beta;
  const rm_handle* row_scale;   /* NULL = None */
  const rm_handle* col_scale;   /* NULL = None */
  rm_scale_op row_op, col_op;
  int has_clamp_min; double clamp_min;
  int has_clamp_max; double clamp_max;
  int has_pow; double pow_exponent;
  const rm_handle* diag_output; /* NULL = None; written in place */
} rm_matmul_epilogue;
rm_status rm_matmul(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);
rm_status rm_matmul_epilogue_apply(rm_provider* p, const rm_handle* a, const rm_handle* b,
                                   const rm_matmul_epilogue* ep, rm_handle* out);
rm_status rm_syrk(rm_provider* p, const rm_handle* a, rm_handle* out); /* A' * A, lib.rs:2383 */
/* hooks of the planner's special fusion kinds */
rm_status rm_matmul_power_step(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, double epsilon, rm_handle* out); /* lib.rs:2414 */
rm_status rm_covariance(rm_provider* p, const rm_handle* matrix, int normalization_biased, rm_handle* out); /* lib.rs:2857, rows=All, unweighted */
rm_status rm_diag_extract(rm_provider* p, const rm_handle* matrix, int64_t offset, rm_handle* out);         /* lib.rs:1626 */
/* selects the GEMM engine: 0 = auto, 1 = FP64 DMMA (mma.sync m8n8k4), 2 = Ozaki split on tcgen05 i8.
 * The tcgen05 engine is element-wise safe on every input: entries whose terms hide under the row/column maximum, and
 * non-finite inputs, are detected on the device and recomputed by the FP64 kernel in the same stream (gemm_ozaki.cu). */
rm_status rm_set_matmul_engine(rm_provider* p, int engine);
/* test/debug: waits for the stream; out4 = {non-finite input seen, pipeline error, tiles recomputed in FP64, int8 GEMMs per tile
 * (digit products + guard)} of the last tcgen05 product */
rm_status rm_debug_ozaki_stats(rm_provider* p, int32_t* out4);
/* measurement / tuning: the generated fused kernels are launched with programmatic stream serialisation (each opens with
 * griddepcontrol.launch_dependents + griddepcontrol.wait, so a kernel becomes resident while its predecessor drains; stream order
 * of all memory effects is unchanged). 0 switches to plain launches, e.g. to time one kernel in isolation. Default: enabled. */
rm_status rm_set_launch_overlap(rm_provider* p, int enabled);
/* test/debug: waits for the stream; device-side protocol flags, all zero unless a bounded pipeline wait ran out
 * ([0] = the TMA-staged imfilter kernel) */
rm_status rm_debug_device_flags(rm_provider* p, int32_t* out, uint32_t n);

/* ---- a9: mldivide core (lib.rs:2477-2489): square systems by device LU with partial pivoting; non-square,
 *      singular or badly conditioned inputs return RM_UNSUPPORTED (host SVD fallback, as with wgpu today) ---- */
rm_status rm_mldivide(rm_provider* p, const rm_handle* a, const rm_handle* b, rm_handle* out);
rm_status rm_mrdivide(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, rm_handle* out); /* lhs / rhs, lib.rs:2484 */
/* linsolve (lib.rs:2422-2429), ProviderLinsolveOptions (lib.rs:681-691), ProviderLinsolveResult (lib.rs:694-697).
 * On the device: LT / UT triangular systems (TRANSA included; rcond = min|diag|/max|diag| like the host's diagonal_rcond,
 * zero diagonal and opts.rcond violations raise the host's "singular to working precision" error) and, when no reciprocal
 * condition number is requested, square general / SYM / POSDEF systems through the LU of rm_mldivide (rcond = NaN).
 * RECT, and general solves that need the singular-value rcond, return RM_UNSUPPORTED (host fallback). */
typedef struct rm_linsolve_options {
  int lower, upper, rectangular, transposed, conjugate, symmetric, posdef, need_rcond;
  int has_rcond; double rcond; /* Option<f64> */
} rm_linsolve_options;
rm_status rm_linsolve(rm_provider* p, const rm_handle* lhs, const rm_handle* rhs, const rm_linsolve_options* opt,
                      rm_handle* solution, double* reciprocal_condition);

/* ---- a10/a11: Monte-Carlo evolution + RNG (lib.rs:1713-1775) -------------------------------------- */
rm_status rm_set_rng_state(rm_provider* p, uint64_t state);                                     /* :1772 */
rm_status rm_get_rng_state(rm_provider* p, uint64_t* state);
rm_status rm_random_uniform(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out); /* :1713 */
rm_status rm_random_normal(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out);  /* :1723 */
rm_status rm_stochastic_evolution(rm_provider* p, const rm_handle* state, double drift, double scale,
                                  uint32_t steps, rm_handle* out);                                /* :1759 */
/* Sharded form: this rank owns paths [path_offset, path_offset+len) of a global vector of
 * `global_len` paths; the RNG stream addressing is that of the single-GPU run, so the union of all
 * ranks' outputs is bit-identical to one rank evolving the whole vector. */
rm_status rm_stochastic_evolution_sharded(rm_provider* p, const rm_handle* state, double drift,
                                          double scale, uint32_t steps, uint64_t path_offset,
                                          uint64_t global_len, rm_handle* out);
/* fused payoff partial: sum_i max(S_i - strike, 0) over this rank's paths -> 1x1 f64 handle that
 * feeds the single NCCL allreduce (benchmarks/monte-carlo-analysis/runmat_lcg.m:48-49). */
rm_status rm_payoff_partial_sum(rm_provider* p, const rm_handle* state, double strike, rm_handle* out);

/* ---- a12/a13: image normalise + imfilter (lib.rs:2407-2413, 3564-3577, 1810-1817) ----------------- */
typedef struct rm_image_normalize_desc {
  uint64_t batch, height, width;
  double epsilon;
  int has_gain; double gain;
  int has_bias; double bias;
  int has_gamma; double gamma;
  int clamp_zero;
} rm_image_normalize_desc;
rm_status rm_image_normalize(rm_provider* p, const rm_handle* input, const rm_image_normalize_desc* d,
                             rm_handle* out);
typedef enum rm_imfilter_padding { RM_PAD_CONSTANT = 0, RM_PAD_REPLICATE, RM_PAD_SYMMETRIC, RM_PAD_CIRCULAR } rm_imfilter_padding;
typedef enum rm_imfilter_shape { RM_IMF_SAME = 0, RM_IMF_FULL = 1, RM_IMF_VALID = 2 } rm_imfilter_shape;
typedef enum rm_imfilter_mode { RM_IMF_CORR = 0, RM_IMF_CONV = 1 } rm_imfilter_mode;
typedef struct rm_imfilter_options {
  rm_imfilter_padding padding;
  double constant_value;
  rm_imfilter_shape shape;
  rm_imfilter_mode mode;
} rm_imfilter_options;
rm_status rm_imfilter(rm_provider* p, const rm_handle* image, const rm_handle* kernel,
                      const rm_imfilter_options* opt, rm_handle* out);
/* conv2d (lib.rs:2543): ProviderConvMode Full=0 / Same=1 / Valid=2; full convolution then MATLAB conv2 slicing
 * (simple_provider.rs:1845-1956, 6065-6155) */
typedef enum rm_conv_mode { RM_CONV_FULL = 0, RM_CONV_SAME = 1, RM_CONV_VALID = 2 } rm_conv_mode;
rm_status rm_conv2d(rm_provider* p, const rm_handle* signal, const rm_handle* kernel, rm_conv_mode mode, rm_handle* out);

/* ---- a15: telemetry / tuning hints (lib.rs:3010-3060) -------------------------------------------- */
rm_status rm_telemetry_snapshot(rm_provider* p, rm_telemetry* out);
/* ProviderTelemetry::kernel_launches (lib.rs:1353, KernelLaunchTelemetry lib.rs:1369-1375): bounded log of the most recent
 * launches, newest last, at most RM_MAX_KERNEL_LAUNCH_EVENTS = 64 like the reference (accelerate/src/telemetry.rs:12, 219-239).
 * Kernel names and shape keys follow the wgpu provider ("fused_elementwise": len/inputs/rank, "fused_reduction":
 * reduce_len/slices/rank, "matmul": m/n/k, "linsolve_triangular", ...); `tuning` holds this backend's choices. */
#define RM_MAX_KERNEL_LAUNCH_EVENTS 64
#define RM_LAUNCH_ATTRS 4
typedef struct rm_kernel_attr { char key[16]; uint64_t value; } rm_kernel_attr;
typedef struct rm_kernel_launch_event {
  char kernel[32];
  rm_precision precision;
  uint32_t n_shape, n_tuning;
  rm_kernel_attr shape[RM_LAUNCH_ATTRS];
  rm_kernel_attr tuning[RM_LAUNCH_ATTRS];
} rm_kernel_launch_event;
/* copies up to `cap` events (oldest first) into `out`; *count = number written */
rm_status rm_kernel_launch_log(rm_provider* p, rm_kernel_launch_event* out, uint32_t cap, uint32_t* count);
/* spawn_handle_concurrency (lib.rs:1400-1402, SpawnHandleConcurrency lib.rs:825-834): handles are ids into a mutex-guarded
 * table and all work is stream-ordered, so tasks may share and mutate handles like the in-process provider
 * (simple_provider.rs:2605): returns RM_SPAWN_SYNCHRONIZED_MUTATION. */
typedef enum rm_spawn_handle_concurrency { RM_SPAWN_IMMUTABLE_SHARE = 0, RM_SPAWN_COPY_ON_WRITE = 1, RM_SPAWN_SYNCHRONIZED_MUTATION = 2, RM_SPAWN_REJECT = 3 } rm_spawn_handle_concurrency;
rm_spawn_handle_concurrency rm_spawn_handle_concurrency_policy(rm_provider* p);
rm_status rm_reset_telemetry(rm_provider* p);

/* ---- measurement helpers (extension; used by bench.py so timing is taken with CUDA events on the
 *      stream the kernels are launched on) -------------------------------------------------------- */
rm_status rm_timer_begin(rm_provider* p);
rm_status rm_timer_end_ms(rm_provider* p, double* elapsed_ms); /* records + synchronises the end event */
rm_status rm_flush_l2(rm_provider* p);                         /* writes a 256 MiB scratch buffer */
rm_status rm_pinned_alloc(size_t bytes, void** out);           /* page-locked host memory for the e2e leg */
rm_status rm_pinned_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RM_ACCEL_H */
