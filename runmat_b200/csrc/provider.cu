// provider.cu — provider lifecycle, buffer table, upload/download/free, constructors, layout + indexing.
//
// Replaces the reference providers' buffer registries: the host provider's `HashMap<u64, Vec<f64>>`
// (crates/runmat-accelerate/src/simple_provider.rs:65-70, 2719-2768) and the wgpu provider's buffer
// residency pool (backend/wgpu/residency.rs:34-127). B200 design: one stream per provider, stream-ordered
// allocation from a CUDA memory pool with an unbounded release threshold (alloc/free are pointer bumps
// after warm-up; 180 GB HBM means we never trim), handles are plain ids.
#include <cstdarg>

#include "common.h"

namespace rm {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}
const char* last_error() { return g_err.c_str(); }
rm_status fail(rm_status code, const char* fmt, ...) {
  char buf[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

rm_status alloc_tensor(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out, void** dptr) {
  RM_REQUIRE(rank <= RM_MAX_RANK, RM_UNSUPPORTED, "tensor rank %u exceeds RM_MAX_RANK=%d", rank, RM_MAX_RANK);
  const uint64_t elems = shape_elems(shape, rank);
  void* ptr = nullptr;
  const size_t bytes = std::max<size_t>(elems * p->elem_size(), 32);  // never a null buffer; 32 B keeps vector tails legal
  cudaError_t e = cudaMallocAsync(&ptr, bytes, p->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? RM_OOM : RM_ERROR, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  const uint64_t id = p->next_id.fetch_add(1, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->buffers[id] = Buffer{ptr, elems};
  }
  p->live_bytes.fetch_add(bytes, std::memory_order_relaxed);
  memset(out, 0, sizeof *out);
  out->buffer_id = id;
  out->device_id = p->device_id;
  out->rank = rank;
  for (uint32_t i = 0; i < rank; ++i) out->shape[i] = shape[i];
  if (dptr) *dptr = ptr;
  return RM_OK;
}

rm_status resolve(rm_provider* p, const rm_handle* h, void** dptr, uint64_t* elems) {
  RM_REQUIRE(h != nullptr, RM_INVALID_ARG, "null handle");
  RM_REQUIRE(h->device_id == p->device_id, RM_INVALID_HANDLE, "handle belongs to device %u, but this provider owns device %u", h->device_id, p->device_id);
  RM_REQUIRE(h->rank <= RM_MAX_RANK, RM_INVALID_ARG, "handle rank %u exceeds RM_MAX_RANK", h->rank);
  Buffer b;
  cudaEvent_t ready = nullptr;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->buffers.find(h->buffer_id);
    if (it == p->buffers.end()) return fail(RM_INVALID_HANDLE, "buffer not found: %llu", (unsigned long long)h->buffer_id);
    b = it->second;
    ready = it->second.ready;
    it->second.ready = nullptr;
    // first use of an uploaded buffer: order the compute stream after its H2D copy (enqueued while the table lock is held, so
    // a concurrent resolve() of the same buffer cannot launch ahead of the wait), then recycle the event
    if (ready) cudaStreamWaitEvent(p->stream, ready, 0);
    // first use of a peer-exchange result: its (lazy) combine goes onto the compute stream now, ahead of the consumer
    if (it->second.p2p_step1) { p2p_enqueue_combine_locked(p, it->second.p2p_step1 - 1, it->second.ptr); it->second.p2p_step1 = 0; }
  }
  if (ready) {
    std::lock_guard<std::mutex> lk(p->ev_mu);
    p->event_pool.push_back(ready);
  }
  const uint64_t n = handle_elems(h);
  RM_REQUIRE(n == b.elems, RM_INVALID_ARG, "handle shape holds %llu elements but buffer %llu stores %llu",
             (unsigned long long)n, (unsigned long long)h->buffer_id, (unsigned long long)b.elems);
  *dptr = b.ptr;
  if (elems) *elems = b.elems;
  return RM_OK;
}

// Caller holds p->scratch_mu from this call until the launch that consumes the scratch has been enqueued.
rm_status ensure_scratch(rm_provider* p, size_t bytes) {
  if (bytes <= p->reduce_scratch_bytes) return RM_OK;
  size_t want = std::max<size_t>(bytes, 1 << 20);
  if (p->reduce_scratch) RM_CUDA(cudaFreeAsync(p->reduce_scratch, p->stream));
  p->reduce_scratch = nullptr;
  p->reduce_scratch_bytes = 0;
  RM_CUDA(cudaMallocAsync(&p->reduce_scratch, want, p->stream));
  p->reduce_scratch_bytes = want;
  return RM_OK;
}

// ---- small kernels ------------------------------------------------------------------------------------------
template <typename T>
__global__ void fill_kernel(T* __restrict__ out, uint64_t n, T value) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = value;
}
template <typename T>
__global__ void eye_kernel(T* __restrict__ out, uint64_t rows, uint64_t cols, uint64_t pages) {
  const uint64_t n = rows * cols * pages;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i % rows, c = (i / rows) % cols;
    out[i] = r == c ? (T)1 : (T)0;
  }
}
// simple_provider.rs:3488-3512: start + idx*step, last element forced to stop (no FMA: built with -fmad=false)
template <typename T>
__global__ void linspace_kernel(T* __restrict__ out, uint64_t count, double start, double stop, double step) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (T)(i == count - 1 ? stop : start + (double)i * step);
}
template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ in, D* __restrict__ out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = (D)in[i];
}
// 32x32 shared-memory tile transpose (padded: no bank conflicts), coalesced on both sides
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t rows, uint64_t cols) {
  __shared__ T tile[32][33];
  const uint64_t r0 = (uint64_t)blockIdx.x * 32, c0 = (uint64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const uint64_t r = r0 + threadIdx.x, c = c0 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[r + c * rows];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const uint64_t c = c0 + threadIdx.x, r = r0 + j;
    if (r < rows && c < cols) out[c + r * cols] = tile[threadIdx.x][j];
  }
}
struct GatherND {
  uint32_t rank;
  uint64_t out_shape[RM_MAX_RANK];
  uint64_t src_stride[RM_MAX_RANK];  // element stride in the source for each OUTPUT dim
  uint64_t src_mod[RM_MAX_RANK];     // 0 = plain, else coordinate is taken modulo this (repmat)
};
template <typename T>
__global__ void gather_nd_kernel(const T* __restrict__ in, T* __restrict__ out, uint64_t n, const __grid_constant__ GatherND g) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t rem = i, src = 0;
    for (uint32_t d = 0; d < g.rank; ++d) {
      uint64_t c = rem % g.out_shape[d];
      rem /= g.out_shape[d];
      if (g.src_mod[d]) c %= g.src_mod[d];
      src += c * g.src_stride[d];
    }
    out[i] = in[src];
  }
}
template <typename T>
__global__ void gather_linear_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = src[idx[i]];
}
template <typename T>
__global__ void scatter_linear_kernel(T* __restrict__ dst, const uint32_t* __restrict__ idx, const T* __restrict__ vals, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[idx[i]] = vals[i];
}

template <typename T>
__global__ void scatter_pos_kernel(T* __restrict__ dst, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ pos, const T* __restrict__ vals, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[idx[i]] = vals[pos[i]];
}

static inline unsigned grid_for(rm_provider* p, uint64_t n, unsigned block = 256) {
  uint64_t b = (n + block - 1) / block;
  uint64_t cap = (uint64_t)p->prop.multiProcessorCount * 16;
  return (unsigned)std::max<uint64_t>(1, std::min(b, cap));
}

#define DISPATCH_T(p, CALL_F64, CALL_F32) \
  do { if ((p)->precision == RM_F64) { CALL_F64; } else { CALL_F32; } } while (0)

static rm_status fill_impl(rm_provider* p, const uint64_t* shape, uint32_t rank, double value, rm_handle* out) {
  void* ptr;
  RM_TRY(alloc_tensor(p, shape, rank, out, &ptr));
  const uint64_t n = shape_elems(shape, rank);
  if (n == 0) return RM_OK;
  if (value == 0.0) {
    RM_CUDA(cudaMemsetAsync(ptr, 0, n * p->elem_size(), p->stream));
    return RM_OK;
  }
  DISPATCH_T(p, (fill_kernel<double><<<grid_for(p, n), 256, 0, p->stream>>>((double*)ptr, n, value)),
             (fill_kernel<float><<<grid_for(p, n), 256, 0, p->stream>>>((float*)ptr, n, (float)value)));
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

}  // namespace rm

using namespace rm;

// =============================================================================================================
// lifecycle
// =============================================================================================================
RM_EXPORT uint32_t rm_abi_version(void) { return RM_ABI_VERSION; }
RM_EXPORT const char* rm_last_error(void) { return rm::last_error(); }

RM_EXPORT rm_status rm_provider_create(int ordinal, uint32_t device_id, rm_precision precision, rm_provider** out) {
  RM_REQUIRE(out != nullptr, RM_INVALID_ARG, "rm_provider_create: null out pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(RM_NO_DEVICE, "no CUDA device available (%s); this backend has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  RM_REQUIRE(ordinal >= 0 && ordinal < count, RM_NO_DEVICE, "CUDA ordinal %d out of range (have %d devices)", ordinal, count);
  RM_CUDA(cudaSetDevice(ordinal));
  std::unique_ptr<rm_provider> p(new rm_provider());
  p->ordinal = ordinal;
  p->device_id = device_id;
  p->precision = precision;
  if (getenv("RUNMAT_B200_NO_PDL")) p->launch_overlap = false;  // A/B: plain launches everywhere (fused.cu and common.h launch_pdl)
  RM_CUDA(cudaGetDeviceProperties(&p->prop, ordinal));
  RM_REQUIRE(p->prop.major >= 10, RM_NO_DEVICE, "device %s is sm_%d%d; this backend is built for sm_100a only", p->prop.name, p->prop.major, p->prop.minor);
  RM_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  RM_CUDA(cudaDeviceGetDefaultMemPool(&p->pool, ordinal));
  uint64_t threshold = UINT64_MAX;
  RM_CUDA(cudaMemPoolSetAttribute(p->pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  RM_CUDA(cudaEventCreate(&p->ev_begin));
  RM_CUDA(cudaEventCreate(&p->ev_end));
  RM_CUDA(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));

  RM_TRY(fused_cache_create(p.get()));
  *out = p.release();
  return RM_OK;
}

RM_EXPORT rm_status rm_provider_destroy(rm_provider* p) {
  if (!p) return RM_OK;
  DeviceGuard g(p->ordinal);
  cudaStreamSynchronize(p->stream);
  if (p->h2d_stream) cudaStreamSynchronize(p->h2d_stream);
  for (auto& kv : p->buffers) { if (kv.second.ready) { cudaEventDestroy(kv.second.ready); kv.second.ready = nullptr; } cudaFreeAsync(kv.second.ptr, p->stream); }
  p->buffers.clear();
  if (p->reduce_scratch) cudaFreeAsync(p->reduce_scratch, p->stream);
  if (p->l2_flush) cudaFreeAsync(p->l2_flush, p->stream);
  if (p->dev_flags) cudaFree(p->dev_flags);
  if (p->aux_stream) { cudaStreamSynchronize(p->aux_stream); cudaStreamDestroy(p->aux_stream); p->aux_stream = nullptr; }
  ozaki_workspace_destroy(p);
  cudaStreamSynchronize(p->stream);
  comm_destroy(p);
  fused_cache_destroy(p);
  if (p->h2d_stream) { cudaStreamSynchronize(p->h2d_stream); cudaStreamDestroy(p->h2d_stream); }
  for (auto& kv : p->buffers) if (kv.second.ready) cudaEventDestroy(kv.second.ready);
  for (cudaEvent_t e : p->event_pool) cudaEventDestroy(e);
  if (p->ev_begin) cudaEventDestroy(p->ev_begin);
  if (p->ev_end) cudaEventDestroy(p->ev_end);
  if (p->owns_stream && p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return RM_OK;
}

RM_EXPORT rm_status rm_device_info_string(rm_provider* p, char* buf, size_t buflen) {
  RM_REQUIRE(p && buf && buflen, RM_INVALID_ARG, "rm_device_info_string: bad arguments");
  snprintf(buf, buflen, "%s (CUDA sm_%d%d, %d SMs, %.0f GiB, runmat-b200 provider, %s)", p->prop.name, p->prop.major, p->prop.minor,
           p->prop.multiProcessorCount, (double)p->prop.totalGlobalMem / (1024.0 * 1024.0 * 1024.0), p->precision == RM_F64 ? "f64" : "f32");
  return RM_OK;
}
RM_EXPORT rm_status rm_device_info_struct(rm_provider* p, rm_device_info* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "rm_device_info_struct: bad arguments");
  memset(out, 0, sizeof *out);
  out->device_id = p->device_id;
  snprintf(out->name, sizeof out->name, "%s", p->prop.name);
  snprintf(out->vendor, sizeof out->vendor, "NVIDIA");
  snprintf(out->backend, sizeof out->backend, "cuda-sm_100a");
  out->memory_bytes = p->prop.totalGlobalMem;
  out->sm_count = (uint32_t)p->prop.multiProcessorCount;
  out->cc_major = (uint32_t)p->prop.major;
  out->cc_minor = (uint32_t)p->prop.minor;
  return RM_OK;
}
RM_EXPORT rm_status rm_device_pci_bus_id(rm_provider* p, char* buf, uint32_t buflen) {
  RM_REQUIRE(p && buf && buflen >= 13, RM_INVALID_ARG, "rm_device_pci_bus_id: bad arguments");
  RM_CUDA(cudaDeviceGetPCIBusId(buf, (int)buflen, p->ordinal));
  return RM_OK;
}
RM_EXPORT uint32_t rm_device_id(rm_provider* p) { return p ? p->device_id : 0; }
RM_EXPORT rm_precision rm_provider_precision(rm_provider* p) { return p ? p->precision : RM_F64; }
RM_EXPORT rm_status rm_synchronize(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  DeviceGuard g(p->ordinal);
  RM_CUDA(cudaStreamSynchronize(p->stream));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  return RM_OK;
}
RM_EXPORT uint64_t rm_host_sync_count(rm_provider* p) { return p ? p->host_syncs.load() : 0; }
RM_EXPORT rm_status rm_debug_ozaki_stats(rm_provider* p, int32_t* out4) {
  RM_REQUIRE(p && out4, RM_INVALID_ARG, "rm_debug_ozaki_stats: bad arguments");
  DeviceGuard g(p->ordinal);
  int tmp[4];
  RM_TRY(ozaki_last_stats(p, tmp));
  for (int i = 0; i < 4; ++i) out4[i] = tmp[i];
  return RM_OK;
}
RM_EXPORT rm_status rm_set_launch_overlap(rm_provider* p, int enabled) {
  RM_REQUIRE(p, RM_INVALID_ARG, "rm_set_launch_overlap: null provider");
  p->launch_overlap = enabled != 0;
  return RM_OK;
}
RM_EXPORT rm_status rm_debug_device_flags(rm_provider* p, int32_t* out, uint32_t n) {
  RM_REQUIRE(p && out && n <= 64, RM_INVALID_ARG, "rm_debug_device_flags: bad arguments");
  for (uint32_t i = 0; i < n; ++i) out[i] = 0;
  if (!p->dev_flags) return RM_OK;
  DeviceGuard g(p->ordinal);
  RM_CUDA(cudaStreamSynchronize(p->stream));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  RM_CUDA(cudaMemcpy(out, p->dev_flags, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return RM_OK;
}
RM_EXPORT rm_status rm_get_stream(rm_provider* p, void** s) {
  RM_REQUIRE(p && s, RM_INVALID_ARG, "rm_get_stream: bad arguments");
  *s = (void*)p->stream;
  return RM_OK;
}
RM_EXPORT rm_status rm_set_stream(rm_provider* p, void* s) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  DeviceGuard g(p->ordinal);
  RM_CUDA(cudaStreamSynchronize(p->stream));
  if (p->owns_stream && p->stream) cudaStreamDestroy(p->stream);
  p->stream = (cudaStream_t)s;
  p->owns_stream = false;
  return RM_OK;
}
RM_EXPORT rm_status rm_device_ptr(rm_provider* p, const rm_handle* h, void** dptr, uint64_t* elems) {
  RM_REQUIRE(p && dptr, RM_INVALID_ARG, "rm_device_ptr: bad arguments");
  return resolve(p, h, dptr, elems);
}
RM_EXPORT rm_status rm_copy_to_device(rm_provider* p, const rm_handle* h, void* dst, uint64_t dst_elems) {
  RM_REQUIRE(p && h && dst, RM_INVALID_ARG, "copy_to_device: bad arguments");
  DeviceGuard g(p->ordinal);
  void* src;
  uint64_t n;
  RM_TRY(resolve(p, h, &src, &n));
  RM_REQUIRE(dst_elems >= n, RM_INVALID_ARG, "copy_to_device: destination holds %llu elements, tensor has %llu", (unsigned long long)dst_elems, (unsigned long long)n);
  RM_CUDA(cudaMemcpyAsync(dst, src, n * p->elem_size(), cudaMemcpyDeviceToDevice, p->stream));
  return RM_OK;
}
RM_EXPORT uint64_t rm_live_buffers(rm_provider* p) { std::lock_guard<std::mutex> lk(p->mu); return p->buffers.size(); }
RM_EXPORT uint64_t rm_live_bytes(rm_provider* p) { return p->live_bytes.load(); }

// =============================================================================================================
// a2: upload / download / free
// =============================================================================================================
template <typename HostT>
static rm_status upload_impl(rm_provider* p, const HostT* data, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && out && (shape || rank == 0), RM_INVALID_ARG, "upload: bad arguments");
  DeviceGuard g(p->ordinal);
  const uint64_t n = shape_elems(shape, rank);
  RM_REQUIRE(data != nullptr || n == 0, RM_INVALID_ARG, "upload: null host data");
  const bool dev_f64 = p->precision == RM_F64;
  const bool host_f64 = sizeof(HostT) == 8;
  void* ptr = nullptr;
  if (dev_f64 == host_f64 && n > 0) {
    // Allocation and copy both live on the dedicated H2D stream, so an upload never queues behind kernels or D2H copies on
    // the compute stream. The buffer carries a "ready" event; resolve() makes the compute stream wait for it at first use.
    RM_REQUIRE(rank <= RM_MAX_RANK, RM_UNSUPPORTED, "tensor rank %u exceeds RM_MAX_RANK=%d", rank, RM_MAX_RANK);
    const size_t bytes = std::max<size_t>(n * sizeof(HostT), 32);
    cudaError_t e = cudaMallocAsync(&ptr, bytes, p->h2d_stream);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? RM_OOM : RM_ERROR, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); }
    cudaEvent_t ev = nullptr;
    {
      std::lock_guard<std::mutex> lk(p->ev_mu);
      if (!p->event_pool.empty()) { ev = p->event_pool.back(); p->event_pool.pop_back(); }
    }
    if (!ev) RM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RM_CUDA(cudaMemcpyAsync(ptr, data, n * sizeof(HostT), cudaMemcpyHostToDevice, p->h2d_stream));
    RM_CUDA(cudaEventRecord(ev, p->h2d_stream));
    const uint64_t id = p->next_id.fetch_add(1, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(p->mu);
      Buffer b;
      b.ptr = ptr;
      b.elems = n;
      b.ready = ev;
      p->buffers[id] = b;
    }
    p->live_bytes.fetch_add(bytes, std::memory_order_relaxed);
    memset(out, 0, sizeof *out);
    out->buffer_id = id;
    out->device_id = p->device_id;
    out->rank = rank;
    for (uint32_t i = 0; i < rank; ++i) out->shape[i] = shape[i];
    p->upload_bytes.fetch_add(n * sizeof(HostT), std::memory_order_relaxed);
    return RM_OK;
  }
  RM_TRY(alloc_tensor(p, shape, rank, out, &ptr));
  if (n == 0) return RM_OK;
  if (dev_f64 == host_f64) {
    RM_CUDA(cudaMemcpyAsync(ptr, data, n * sizeof(HostT), cudaMemcpyHostToDevice, p->stream));
  } else {
    // precision mismatch: stage the host bytes, convert on the device (wgpu narrows per element: io.rs:87)
    void* stage = nullptr;
    RM_CUDA(cudaMallocAsync(&stage, n * sizeof(HostT), p->stream));
    RM_CUDA(cudaMemcpyAsync(stage, data, n * sizeof(HostT), cudaMemcpyHostToDevice, p->stream));
    if (host_f64) convert_kernel<double, float><<<grid_for(p, n), 256, 0, p->stream>>>((const double*)stage, (float*)ptr, n);
    else convert_kernel<float, double><<<grid_for(p, n), 256, 0, p->stream>>>((const float*)stage, (double*)ptr, n);
    RM_LAUNCH_CHECK();
    count_launch(p);
    RM_CUDA(cudaFreeAsync(stage, p->stream));
  }
  p->upload_bytes.fetch_add(n * sizeof(HostT), std::memory_order_relaxed);
  return RM_OK;
}
RM_EXPORT rm_status rm_upload(rm_provider* p, const double* data, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  return upload_impl<double>(p, data, shape, rank, out);
}
RM_EXPORT rm_status rm_upload_f32(rm_provider* p, const float* data, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  return upload_impl<float>(p, data, shape, rank, out);
}

template <typename HostT>
static rm_status download_impl(rm_provider* p, const rm_handle* h, HostT* out, uint64_t out_len) {
  RM_REQUIRE(p && h, RM_INVALID_ARG, "download: bad arguments");
  DeviceGuard g(p->ordinal);
  void* ptr;
  uint64_t n;
  RM_TRY(resolve(p, h, &ptr, &n));
  RM_REQUIRE(out_len >= n, RM_INVALID_ARG, "download: host buffer holds %llu elements, tensor has %llu", (unsigned long long)out_len, (unsigned long long)n);
  if (n == 0) return RM_OK;
  RM_REQUIRE(out != nullptr, RM_INVALID_ARG, "download: null host buffer");
  const bool dev_f64 = p->precision == RM_F64;
  const bool host_f64 = sizeof(HostT) == 8;
  if (dev_f64 == host_f64) {
    RM_CUDA(cudaMemcpyAsync(out, ptr, n * sizeof(HostT), cudaMemcpyDeviceToHost, p->stream));
  } else {
    void* stage = nullptr;
    RM_CUDA(cudaMallocAsync(&stage, n * sizeof(HostT), p->stream));
    if (host_f64) convert_kernel<float, double><<<grid_for(p, n), 256, 0, p->stream>>>((const float*)ptr, (double*)stage, n);
    else convert_kernel<double, float><<<grid_for(p, n), 256, 0, p->stream>>>((const double*)ptr, (float*)stage, n);
    RM_LAUNCH_CHECK();
    count_launch(p);
    RM_CUDA(cudaMemcpyAsync(out, stage, n * sizeof(HostT), cudaMemcpyDeviceToHost, p->stream));
    RM_CUDA(cudaFreeAsync(stage, p->stream));
  }
  RM_CUDA(cudaStreamSynchronize(p->stream));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  p->download_bytes.fetch_add(n * sizeof(HostT), std::memory_order_relaxed);
  return RM_OK;
}
RM_EXPORT rm_status rm_download(rm_provider* p, const rm_handle* h, double* out, uint64_t out_len) { return download_impl<double>(p, h, out, out_len); }
RM_EXPORT rm_status rm_download_f32(rm_provider* p, const rm_handle* h, float* out, uint64_t out_len) { return download_impl<float>(p, h, out, out_len); }

// Stream-ordered download without the final synchronisation: the host buffer is valid after rm_synchronize(). Lets a
// caller overlap the D2H of one chunk with the upload/compute of the next (bench.py e2e leg).
RM_EXPORT rm_status rm_download_async(rm_provider* p, const rm_handle* h, double* out, uint64_t out_len) {
  RM_REQUIRE(p && h && out, RM_INVALID_ARG, "download_async: bad arguments");
  RM_REQUIRE(p->precision == RM_F64, RM_UNSUPPORTED, "download_async: f32 storage not supported (use rm_download)");
  DeviceGuard g(p->ordinal);
  void* ptr;
  uint64_t n;
  RM_TRY(resolve(p, h, &ptr, &n));
  RM_REQUIRE(out_len >= n, RM_INVALID_ARG, "download_async: host buffer holds %llu elements, tensor has %llu", (unsigned long long)out_len, (unsigned long long)n);
  if (n) RM_CUDA(cudaMemcpyAsync(out, ptr, n * 8, cudaMemcpyDeviceToHost, p->stream));
  p->download_bytes.fetch_add(n * 8, std::memory_order_relaxed);
  return RM_OK;
}

RM_EXPORT rm_status rm_free(rm_provider* p, const rm_handle* h) {
  RM_REQUIRE(p && h, RM_INVALID_ARG, "free: bad arguments");
  // simple_provider.rs:2752-2768: reject foreign device ids; unknown ids are a no-op
  RM_REQUIRE(h->device_id == p->device_id, RM_INVALID_HANDLE, "free: handle belongs to device %u, but this provider owns device %u", h->device_id, p->device_id);
  Buffer b;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->buffers.find(h->buffer_id);
    if (it == p->buffers.end()) return RM_OK;
    b = it->second;
    if (b.p2p_step1) {
      // freed before use: the combine still runs (it is also this rank's flow-control step of the exchange protocol, comm.cu)
      DeviceGuard g2(p->ordinal);
      p2p_enqueue_combine_locked(p, b.p2p_step1 - 1, b.ptr);
    }
    p->buffers.erase(it);
  }
  DeviceGuard g(p->ordinal);
  p->live_bytes.fetch_sub(std::max<size_t>(b.elems * p->elem_size(), 32), std::memory_order_relaxed);
  if (b.ready) {  // uploaded but never used: the free must still be ordered after the copy
    cudaStreamWaitEvent(p->stream, b.ready, 0);
    std::lock_guard<std::mutex> lk(p->ev_mu);
    p->event_pool.push_back(b.ready);
  }
  RM_CUDA(cudaFreeAsync(b.ptr, p->stream));
  return RM_OK;
}

RM_EXPORT rm_status rm_read_scalar(rm_provider* p, const rm_handle* h, uint64_t linear_index, double* out) {
  RM_REQUIRE(p && h && out, RM_INVALID_ARG, "read_scalar: bad arguments");
  DeviceGuard g(p->ordinal);
  void* ptr;
  uint64_t n;
  RM_TRY(resolve(p, h, &ptr, &n));
  RM_REQUIRE(linear_index < n, RM_INVALID_ARG, "read_scalar: index %llu out of bounds (%llu elements)", (unsigned long long)linear_index, (unsigned long long)n);
  if (p->precision == RM_F64) {
    RM_CUDA(cudaMemcpyAsync(out, (double*)ptr + linear_index, 8, cudaMemcpyDeviceToHost, p->stream));
    RM_CUDA(cudaStreamSynchronize(p->stream));
  } else {
    float f;
    RM_CUDA(cudaMemcpyAsync(&f, (float*)ptr + linear_index, 4, cudaMemcpyDeviceToHost, p->stream));
    RM_CUDA(cudaStreamSynchronize(p->stream));
    *out = (double)f;
  }
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  p->download_bytes.fetch_add(p->elem_size(), std::memory_order_relaxed);
  return RM_OK;
}

// =============================================================================================================
// a14: constructors, reshape, layout, indexing
// =============================================================================================================
RM_EXPORT rm_status rm_fill(rm_provider* p, const uint64_t* shape, uint32_t rank, double value, rm_handle* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "fill: bad arguments");
  DeviceGuard g(p->ordinal);
  return fill_impl(p, shape, rank, value, out);
}
RM_EXPORT rm_status rm_zeros(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) { return rm_fill(p, shape, rank, 0.0, out); }
RM_EXPORT rm_status rm_ones(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) { return rm_fill(p, shape, rank, 1.0, out); }

RM_EXPORT rm_status rm_eye(rm_provider* p, const uint64_t* shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && out && rank >= 1, RM_INVALID_ARG, "eye: bad arguments");
  DeviceGuard g(p->ordinal);
  // simple_provider.rs:3471-3482: identity on the leading two dims, replicated over trailing pages
  uint64_t shp[RM_MAX_RANK];
  uint32_t rk = rank;
  for (uint32_t i = 0; i < rank; ++i) shp[i] = shape[i];
  if (rank == 1) { shp[1] = shape[0]; rk = 2; }
  void* ptr;
  RM_TRY(alloc_tensor(p, shp, rk, out, &ptr));
  const uint64_t n = shape_elems(shp, rk);
  if (n == 0) return RM_OK;
  const uint64_t pages = n / (shp[0] * shp[1]);
  DISPATCH_T(p, (eye_kernel<double><<<grid_for(p, n), 256, 0, p->stream>>>((double*)ptr, shp[0], shp[1], pages)),
             (eye_kernel<float><<<grid_for(p, n), 256, 0, p->stream>>>((float*)ptr, shp[0], shp[1], pages)));
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

RM_EXPORT rm_status rm_linspace(rm_provider* p, double start, double stop, uint64_t count, rm_handle* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "linspace: bad arguments");
  DeviceGuard g(p->ordinal);
  uint64_t shp[2] = {1, count};
  void* ptr;
  RM_TRY(alloc_tensor(p, shp, 2, out, &ptr));
  if (count == 0) return RM_OK;
  const double step = count > 1 ? (stop - start) / (double)(count - 1) : 0.0;
  DISPATCH_T(p, (linspace_kernel<double><<<grid_for(p, count), 256, 0, p->stream>>>((double*)ptr, count, start, stop, step)),
             (linspace_kernel<float><<<grid_for(p, count), 256, 0, p->stream>>>((float*)ptr, count, start, stop, step)));
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

RM_EXPORT rm_status rm_reshape(rm_provider* p, const rm_handle* h, const uint64_t* new_shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && h && out && rank <= RM_MAX_RANK, RM_INVALID_ARG, "reshape: bad arguments");
  // trait default (lib.rs:2676-2684): metadata only. We additionally reject element-count changes.
  void* ptr;
  uint64_t n;
  RM_TRY(resolve(p, h, &ptr, &n));
  RM_REQUIRE(shape_elems(new_shape, rank) == n, RM_INVALID_ARG, "reshape: cannot reshape %llu elements", (unsigned long long)n);
  rm_handle r = *h;
  r.rank = rank;
  memset(r.shape, 0, sizeof r.shape);
  for (uint32_t i = 0; i < rank; ++i) r.shape[i] = new_shape[i];
  *out = r;
  return RM_OK;
}

RM_EXPORT rm_status rm_transpose(rm_provider* p, const rm_handle* a, rm_handle* out) {
  RM_REQUIRE(p && a && out, RM_INVALID_ARG, "transpose: bad arguments");
  RM_REQUIRE(a->rank == 2, RM_ERROR, "transpose: only 2D tensors supported");  // simple_provider.rs:5983-
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, a, &src, nullptr));
  const uint64_t rows = a->shape[0], cols = a->shape[1];
  uint64_t shp[2] = {cols, rows};
  void* dst;
  RM_TRY(alloc_tensor(p, shp, 2, out, &dst));
  if (rows * cols == 0) return RM_OK;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
  RM_REQUIRE(grid.y <= 65535, RM_UNSUPPORTED, "transpose: too many columns");
  DISPATCH_T(p, (transpose_kernel<double><<<grid, block, 0, p->stream>>>((const double*)src, (double*)dst, rows, cols)),
             (transpose_kernel<float><<<grid, block, 0, p->stream>>>((const float*)src, (float*)dst, rows, cols)));
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

static rm_status gather_nd(rm_provider* p, const void* src, const GatherND& g, const uint64_t* out_shape, uint32_t rank, rm_handle* out) {
  void* dst;
  RM_TRY(alloc_tensor(p, out_shape, rank, out, &dst));
  const uint64_t n = shape_elems(out_shape, rank);
  if (n == 0) return RM_OK;
  DISPATCH_T(p, (gather_nd_kernel<double><<<grid_for(p, n), 256, 0, p->stream>>>((const double*)src, (double*)dst, n, g)),
             (gather_nd_kernel<float><<<grid_for(p, n), 256, 0, p->stream>>>((const float*)src, (float*)dst, n, g)));
  RM_LAUNCH_CHECK();
  count_launch(p);
  return RM_OK;
}

RM_EXPORT rm_status rm_permute(rm_provider* p, const rm_handle* a, const uint32_t* order, uint32_t n_order, rm_handle* out) {
  RM_REQUIRE(p && a && order && out, RM_INVALID_ARG, "permute: bad arguments");
  RM_REQUIRE(n_order >= a->rank && n_order <= RM_MAX_RANK, RM_INVALID_ARG, "permute: order length %u must cover rank %u", n_order, a->rank);
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, a, &src, nullptr));
  // order is zero-based: output dim d takes source dim order[d]
  uint64_t ext[RM_MAX_RANK], stride[RM_MAX_RANK];
  uint64_t s = 1;
  for (uint32_t d = 0; d < n_order; ++d) { ext[d] = d < a->rank ? a->shape[d] : 1; stride[d] = s; s *= ext[d]; }
  bool seen[RM_MAX_RANK] = {false};
  GatherND gd{};
  gd.rank = n_order;
  uint64_t oshape[RM_MAX_RANK];
  for (uint32_t d = 0; d < n_order; ++d) {
    RM_REQUIRE(order[d] < n_order && !seen[order[d]], RM_INVALID_ARG, "permute: order is not a permutation");
    seen[order[d]] = true;
    oshape[d] = ext[order[d]];
    gd.out_shape[d] = oshape[d];
    gd.src_stride[d] = stride[order[d]];
    gd.src_mod[d] = 0;
  }
  return gather_nd(p, src, gd, oshape, n_order, out);
}

RM_EXPORT rm_status rm_repmat(rm_provider* p, const rm_handle* a, const uint64_t* reps, uint32_t n_reps, rm_handle* out) {
  RM_REQUIRE(p && a && reps && out, RM_INVALID_ARG, "repmat: bad arguments");
  const uint32_t rank = std::max(a->rank, n_reps);
  RM_REQUIRE(rank <= RM_MAX_RANK, RM_UNSUPPORTED, "repmat: rank too large");
  DeviceGuard g(p->ordinal);
  void* src;
  RM_TRY(resolve(p, a, &src, nullptr));
  GatherND gd{};
  gd.rank = rank;
  uint64_t oshape[RM_MAX_RANK], s = 1;
  for (uint32_t d = 0; d < rank; ++d) {
    const uint64_t ext = d < a->rank ? a->shape[d] : 1, r = d < n_reps ? reps[d] : 1;
    oshape[d] = ext * r;
    gd.out_shape[d] = oshape[d];
    gd.src_stride[d] = s;
    gd.src_mod[d] = ext;
    s *= ext;
  }
  return gather_nd(p, src, gd, oshape, rank, out);
}

RM_EXPORT rm_status rm_cat(rm_provider* p, uint32_t dim_one_based, const rm_handle* inputs, uint32_t n_inputs, rm_handle* out) {
  RM_REQUIRE(p && inputs && out && n_inputs > 0, RM_INVALID_ARG, "cat: bad arguments");
  RM_REQUIRE(dim_one_based >= 1 && dim_one_based <= RM_MAX_RANK, RM_ERROR, "cat: dimension must be >= 1");
  DeviceGuard g(p->ordinal);
  const uint32_t d = dim_one_based - 1;
  uint32_t rank = std::max<uint32_t>(d + 1, 2);
  for (uint32_t i = 0; i < n_inputs; ++i) rank = std::max(rank, inputs[i].rank);
  auto ext = [&](const rm_handle& h, uint32_t k) -> uint64_t { return k < h.rank ? h.shape[k] : 1; };
  uint64_t oshape[RM_MAX_RANK], along = 0;
  for (uint32_t k = 0; k < rank; ++k) oshape[k] = ext(inputs[0], k);
  for (uint32_t i = 0; i < n_inputs; ++i) {
    for (uint32_t k = 0; k < rank; ++k)
      RM_REQUIRE(k == d || ext(inputs[i], k) == oshape[k], RM_ERROR, "cat: dimension mismatch on input %u (dim %u: %llu vs %llu)", i + 1, k + 1,
                 (unsigned long long)ext(inputs[i], k), (unsigned long long)oshape[k]);
    along += ext(inputs[i], d);
  }
  oshape[d] = along;
  void* dst;
  RM_TRY(alloc_tensor(p, oshape, rank, out, &dst));
  uint64_t pre = 1, post = 1;
  for (uint32_t k = 0; k < d; ++k) pre *= oshape[k];
  for (uint32_t k = d + 1; k < rank; ++k) post *= oshape[k];
  const size_t es = p->elem_size();
  uint64_t off = 0;
  for (uint32_t i = 0; i < n_inputs; ++i) {
    void* src;
    rm_status st = resolve(p, &inputs[i], &src, nullptr);
    if (st != RM_OK) { std::string m = last_error(); rm_free(p, out); set_error("%s", m.c_str()); return st; }
    const uint64_t ni = ext(inputs[i], d);
    if (pre * ni * post == 0) continue;
    // each input is `post` slabs of pre*ni contiguous elements; in the output the slabs are pre*along apart
    cudaError_t e = cudaMemcpy2DAsync((char*)dst + off * pre * es, pre * along * es, src, pre * ni * es, pre * ni * es, post, cudaMemcpyDeviceToDevice, p->stream);
    if (e != cudaSuccess) { cudaGetLastError(); rm_free(p, out); return fail(RM_ERROR, "cat: copy failed: %s", cudaGetErrorString(e)); }
    off += ni;
  }
  return RM_OK;
}

static rm_status upload_indices(rm_provider* p, const uint32_t* idx, uint64_t n, uint64_t bound, const char* what, uint32_t** dptr) {
  for (uint64_t i = 0; i < n; ++i)  // bounds are checked on the host, as the reference does (simple_provider.rs:2636-2646)
    RM_REQUIRE(idx[i] < bound, RM_INVALID_ARG, "%s: index %u (position %llu) out of bounds (logical_len=%llu)", what, idx[i], (unsigned long long)i, (unsigned long long)bound);
  RM_CUDA(cudaMallocAsync((void**)dptr, std::max<uint64_t>(n, 1) * 4, p->stream));
  if (n) RM_CUDA(cudaMemcpyAsync(*dptr, idx, n * 4, cudaMemcpyHostToDevice, p->stream));
  return RM_OK;
}

RM_EXPORT rm_status rm_gather_linear(rm_provider* p, const rm_handle* source, const uint32_t* indices, uint64_t n, const uint64_t* out_shape, uint32_t out_rank, rm_handle* out) {
  RM_REQUIRE(p && source && out && (indices || n == 0), RM_INVALID_ARG, "gather_linear: bad arguments");
  DeviceGuard g(p->ordinal);
  void* src;
  uint64_t len;
  RM_TRY(resolve(p, source, &src, &len));
  RM_REQUIRE(shape_elems(out_shape, out_rank) == n, RM_INVALID_ARG, "gather_linear: output shape does not match index count %llu", (unsigned long long)n);
  uint32_t* didx = nullptr;
  RM_TRY(upload_indices(p, indices, n, len, "gather_linear", &didx));
  void* dst;
  rm_status st = alloc_tensor(p, out_shape, out_rank, out, &dst);
  if (st == RM_OK && n) {
    DISPATCH_T(p, (gather_linear_kernel<double><<<grid_for(p, n), 256, 0, p->stream>>>((const double*)src, didx, (double*)dst, n)),
               (gather_linear_kernel<float><<<grid_for(p, n), 256, 0, p->stream>>>((const float*)src, didx, (float*)dst, n)));
    count_launch(p);
  }
  cudaFreeAsync(didx, p->stream);
  RM_LAUNCH_CHECK();
  return st;
}

RM_EXPORT rm_status rm_scatter_linear(rm_provider* p, const rm_handle* target, const uint32_t* indices, uint64_t n, const rm_handle* values) {
  RM_REQUIRE(p && target && values && (indices || n == 0), RM_INVALID_ARG, "scatter_linear: bad arguments");
  DeviceGuard g(p->ordinal);
  void *dst, *vals;
  uint64_t len, vlen;
  RM_TRY(resolve(p, target, &dst, &len));
  RM_TRY(resolve(p, values, &vals, &vlen));
  RM_REQUIRE(vlen == n, RM_INVALID_ARG, "scatter_linear: values raw length %llu does not match index count %llu", (unsigned long long)vlen, (unsigned long long)n);
  // duplicates: the host loop lets the LAST position win (simple_provider.rs:2699-2711). A parallel scatter cannot
  // order writes, so when duplicates exist only the last occurrence of each index is kept (host-side filter).
  std::vector<uint32_t> fidx;
  std::vector<uint32_t> fpos;
  {
    std::unordered_map<uint32_t, uint64_t> last;
    last.reserve(n * 2);
    bool dup = false;
    for (uint64_t i = 0; i < n; ++i) {
      auto it = last.find(indices[i]);
      if (it != last.end()) { dup = true; it->second = i; } else last.emplace(indices[i], i);
    }
    if (dup) {
      RM_REQUIRE(n < (1ull << 32), RM_UNSUPPORTED, "scatter_linear: too many indices");
      for (uint64_t i = 0; i < n; ++i)
        if (last[indices[i]] == i) { fidx.push_back(indices[i]); fpos.push_back((uint32_t)i); }
    }
  }
  uint32_t* didx = nullptr;
  if (fidx.empty()) {
    RM_TRY(upload_indices(p, indices, n, len, "scatter_linear", &didx));
    if (n) {
      DISPATCH_T(p, (scatter_linear_kernel<double><<<grid_for(p, n), 256, 0, p->stream>>>((double*)dst, didx, (const double*)vals, n)),
                 (scatter_linear_kernel<float><<<grid_for(p, n), 256, 0, p->stream>>>((float*)dst, didx, (const float*)vals, n)));
      count_launch(p);
    }
    cudaFreeAsync(didx, p->stream);
  } else {
    const uint64_t m = fidx.size();
    uint32_t* dpos = nullptr;
    RM_TRY(upload_indices(p, fidx.data(), m, len, "scatter_linear", &didx));
    rm_status st = upload_indices(p, fpos.data(), m, n, "scatter_linear", &dpos);
    if (st != RM_OK) { cudaFreeAsync(didx, p->stream); return st; }
    DISPATCH_T(p, (scatter_pos_kernel<double><<<grid_for(p, m), 256, 0, p->stream>>>((double*)dst, didx, dpos, (const double*)vals, m)),
               (scatter_pos_kernel<float><<<grid_for(p, m), 256, 0, p->stream>>>((float*)dst, didx, dpos, (const float*)vals, m)));
    count_launch(p);
    cudaFreeAsync(didx, p->stream);
    cudaFreeAsync(dpos, p->stream);
  }
  RM_LAUNCH_CHECK();
  return RM_OK;
}

// =============================================================================================================
// a15: telemetry + measurement helpers
// =============================================================================================================
RM_EXPORT rm_status rm_telemetry_snapshot(rm_provider* p, rm_telemetry* out) {
  RM_REQUIRE(p && out, RM_INVALID_ARG, "telemetry_snapshot: bad arguments");
  auto snap = [](DispatchCounter& c) { return rm_dispatch_stats{c.count.load(), c.wall_ns.load()}; };
  out->fused_elementwise = snap(p->t_fused_elementwise);
  out->fused_reduction = snap(p->t_fused_reduction);
  out->matmul = snap(p->t_matmul);
  out->linsolve = snap(p->t_linsolve);
  out->mldivide = snap(p->t_mldivide);
  out->mrdivide = snap(p->t_mrdivide);
  out->upload_bytes = p->upload_bytes.load();
  out->download_bytes = p->download_bytes.load();
  out->fusion_cache_hits = p->cache_hits.load();
  out->fusion_cache_misses = p->cache_misses.load();
  out->kernel_launches = p->kernel_launches.load();
  return RM_OK;
}
RM_EXPORT rm_status rm_kernel_launch_log(rm_provider* p, rm_kernel_launch_event* out, uint32_t cap, uint32_t* count) {
  RM_REQUIRE(p && count && (out || cap == 0), RM_INVALID_ARG, "kernel_launch_log: bad arguments");
  std::lock_guard<std::mutex> lk(p->log_mu);
  const uint64_t have = std::min<uint64_t>(p->launch_log_n, RM_MAX_KERNEL_LAUNCH_EVENTS);
  const uint64_t n = std::min<uint64_t>(have, cap);
  for (uint64_t i = 0; i < n; ++i) out[i] = p->launch_log[(p->launch_log_n - n + i) % RM_MAX_KERNEL_LAUNCH_EVENTS];  // the newest n, oldest first
  *count = (uint32_t)n;
  return RM_OK;
}
RM_EXPORT rm_spawn_handle_concurrency rm_spawn_handle_concurrency_policy(rm_provider*) { return RM_SPAWN_SYNCHRONIZED_MUTATION; }
RM_EXPORT rm_status rm_reset_telemetry(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  p->t_fused_elementwise.reset(); p->t_fused_reduction.reset(); p->t_matmul.reset();
  p->t_linsolve.reset(); p->t_mldivide.reset(); p->t_mrdivide.reset();
  p->upload_bytes = 0; p->download_bytes = 0; p->cache_hits = 0; p->cache_misses = 0; p->kernel_launches = 0;
  { std::lock_guard<std::mutex> lk(p->log_mu); p->launch_log_n = 0; }
  return RM_OK;
}
RM_EXPORT void rm_fused_cache_counters(rm_provider* p, uint64_t* hits, uint64_t* misses) {
  if (hits) *hits = p ? p->cache_hits.load() : 0;
  if (misses) *misses = p ? p->cache_misses.load() : 0;
}
RM_EXPORT uint32_t rm_default_reduction_workgroup_size(rm_provider*) { return 256; }
RM_EXPORT uint64_t rm_two_pass_threshold(rm_provider* p) {
  // reductions switch to the two-stage (multi-block + last-block finish) form above this many elements/slice
  return p ? (uint64_t)p->prop.multiProcessorCount * 0 + 8192 : 8192;
}

RM_EXPORT rm_status rm_timer_begin(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  DeviceGuard g(p->ordinal);
  RM_CUDA(cudaEventRecord(p->ev_begin, p->stream));
  return RM_OK;
}
RM_EXPORT rm_status rm_timer_end_ms(rm_provider* p, double* ms) {
  RM_REQUIRE(p && ms, RM_INVALID_ARG, "timer_end: bad arguments");
  DeviceGuard g(p->ordinal);
  RM_CUDA(cudaEventRecord(p->ev_end, p->stream));
  RM_CUDA(cudaEventSynchronize(p->ev_end));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  float f = 0;
  RM_CUDA(cudaEventElapsedTime(&f, p->ev_begin, p->ev_end));
  *ms = (double)f;
  return RM_OK;
}
RM_EXPORT rm_status rm_flush_l2(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "null provider");
  DeviceGuard g(p->ordinal);
  const size_t bytes = 256ull << 20;  // > 126 MB L2
  std::lock_guard<std::mutex> lk(p->scratch_mu);
  if (!p->l2_flush) {
    RM_CUDA(cudaMallocAsync(&p->l2_flush, bytes, p->stream));
    p->l2_flush_bytes = bytes;
  }
  RM_CUDA(cudaMemsetAsync(p->l2_flush, 0, bytes, p->stream));
  return RM_OK;
}
RM_EXPORT rm_status rm_pinned_alloc(size_t bytes, void** out) {
  RM_REQUIRE(out, RM_INVALID_ARG, "pinned_alloc: null out");
  RM_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return RM_OK;
}
RM_EXPORT rm_status rm_pinned_free(void* ptr) {
  if (ptr) RM_CUDA(cudaFreeHost(ptr));
  return RM_OK;
}
