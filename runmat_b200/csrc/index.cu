// index.cu — the indexing class of the "next" rows: find, sub2ind, ind2sub, scatter_column/row.
// All results are bit-exact index arithmetic. Reference semantics: crates/runmat-accelerate/src/simple_provider.rs
//   find            :7500-7573   (1-based linear / row / col of non-zeros, First|Last with optional limit)
//   sub2ind         :8340-8420   (+ coerce_sub2ind_value :2268-2291: finite, integer, within [1, dim])
//   ind2sub         (same validation on the linear index; trait lib.rs:3102-3112)
//   scatter_column / scatter_row  (trait lib.rs:3064-3082: NEW handle with one column / row replaced)
#include "common.h"

namespace rm {

namespace {

constexpr uint64_t SEG = 2048;  // elements per warp segment in the ordered compaction

// pass 1: non-zeros per warp segment (ballot + popc keeps the original order for pass 2)
template <typename T>
__global__ void find_count_kernel(const T* __restrict__ a, uint64_t n, uint64_t* __restrict__ counts, uint64_t nseg) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nseg) return;
  const uint64_t s0 = w * SEG, s1 = min(n, s0 + SEG);
  uint32_t c = 0;
  for (uint64_t i = s0 + lane; i < s0 + SEG; i += 32) {
    const bool nz = i < s1 && a[i] != (T)0;  // NaN != 0 is true, like the host's `value != 0.0`
    c += __popc(__ballot_sync(0xffffffffu, nz));
  }
  if (lane == 0) counts[w] = c;
}
// exclusive scan of the per-segment counts (single CTA; nseg <= a few million)
__global__ void __launch_bounds__(1024) scan_kernel(const uint64_t* __restrict__ counts, uint64_t* __restrict__ offsets, uint64_t nseg, uint64_t* __restrict__ total) {
  __shared__ uint64_t part[1024];
  const uint64_t per = (nseg + 1023) / 1024;
  const uint64_t b = (uint64_t)threadIdx.x * per, e = min(nseg, b + per);
  uint64_t s = 0;
  for (uint64_t i = b; i < e; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t run = 0;
    for (int i = 0; i < 1024; ++i) { const uint64_t v = part[i]; part[i] = run; run += v; }
    *total = run;
  }
  __syncthreads();
  uint64_t run = part[threadIdx.x];
  for (uint64_t i = b; i < e; ++i) { offsets[i] = run; run += counts[i]; }
}
// pass 2: ordered compaction of the zero-based linear indices
template <typename T>
__global__ void find_compact_kernel(const T* __restrict__ a, uint64_t n, const uint64_t* __restrict__ offsets, uint64_t nseg, uint64_t* __restrict__ idx_all) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nseg) return;
  const uint64_t s0 = w * SEG, s1 = min(n, s0 + SEG);
  uint64_t run = offsets[w];
  for (uint64_t i = s0 + lane; i < s0 + SEG; i += 32) {
    const bool nz = i < s1 && a[i] != (T)0;
    const uint32_t mask = __ballot_sync(0xffffffffu, nz);
    if (nz) idx_all[run + __popc(mask & ((1u << lane) - 1u))] = i;
    run += __popc(mask);
  }
}
template <typename T>
__global__ void find_emit_kernel(const T* __restrict__ a, const uint64_t* __restrict__ idx_all, uint64_t found, uint64_t cnt, int last, uint64_t row_extent,
                                 T* __restrict__ linear, T* __restrict__ rows, T* __restrict__ cols, T* __restrict__ values) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t src = last ? idx_all[found - 1 - k] : idx_all[k];
    linear[k] = (T)(src + 1);
    rows[k] = (T)(src % row_extent + 1);
    cols[k] = (T)(src / row_extent + 1);
    values[k] = a[src];
  }
}

struct SubParams {
  uint32_t ndims;
  uint64_t dims[RM_MAX_RANK];
  uint64_t strides[RM_MAX_RANK];
  const void* ptr[RM_MAX_RANK];
  uint8_t scalar[RM_MAX_RANK];
};
// error word: (code << 8) | dim_number ; code 1 non-finite, 2 non-integer, 3 out of range
template <typename T>
__global__ void sub2ind_kernel(const __grid_constant__ SubParams sp, uint64_t len, T* __restrict__ out, int* __restrict__ err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t off = 0;
    for (uint32_t d = 0; d < sp.ndims; ++d) {
      const double raw = (double)(sp.scalar[d] ? ((const T*)sp.ptr[d])[0] : ((const T*)sp.ptr[d])[i]);
      int code = 0;
      const double r = round(raw);
      if (!isfinite(raw)) code = 1;
      else if (fabs(r - raw) > 2.220446049250313e-16) code = 2;
      else if (r < 1.0 || r > (double)sp.dims[d]) code = 3;
      if (code) { atomicCAS(err, 0, (code << 8) | (int)(d + 1)); return; }
      off += ((uint64_t)r - 1) * sp.strides[d];
    }
    out[i] = (T)(off + 1);
  }
}
struct IndParams {
  uint32_t ndims;
  uint64_t dims[RM_MAX_RANK];
  uint64_t strides[RM_MAX_RANK];
  void* out[RM_MAX_RANK];
};
template <typename T>
__global__ void ind2sub_kernel(const T* __restrict__ idx, uint64_t len, uint64_t total, const __grid_constant__ IndParams ip, int* __restrict__ err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) {
    const double raw = (double)idx[i];
    const double r = round(raw);
    int code = 0;
    if (!isfinite(raw)) code = 1;
    else if (fabs(r - raw) > 2.220446049250313e-16) code = 2;
    else if (r < 1.0 || r > (double)total) code = 3;
    if (code) { atomicCAS(err, 0, code << 8); return; }
    const uint64_t z = (uint64_t)r - 1;
    for (uint32_t d = 0; d < ip.ndims; ++d) ((T*)ip.out[d])[i] = (T)((z / ip.strides[d]) % ip.dims[d] + 1);
  }
}
template <typename T>
__global__ void scatter_row_kernel(T* __restrict__ m, uint64_t rows, uint64_t cols, uint64_t row, const T* __restrict__ vals) {
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (uint64_t)gridDim.x * blockDim.x) m[row + c * rows] = vals[c];
}

unsigned grid1d(rm_provider* p, uint64_t n) { return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)p->prop.multiProcessorCount * 16)); }

template <typename T>
rm_status find_impl(rm_provider* p, const rm_handle* a, int has_limit, uint64_t limit, int last, rm_handle* linear, rm_handle* rows, rm_handle* cols, rm_handle* values) {
  void* src;
  uint64_t n;
  RM_TRY(resolve(p, a, &src, &n));
  const uint64_t row_extent = std::max<uint64_t>(a->rank >= 1 ? a->shape[0] : 1, 1);
  uint64_t cap = has_limit ? limit : (last ? 1 : n);  // simple_provider.rs:7514-7518
  cap = std::min(cap, n);
  uint64_t found = 0;
  uint64_t *counts = nullptr, *offsets = nullptr, *idx_all = nullptr, *dtotal = nullptr;
  cudaStream_t st = p->stream;
  auto cleanup = [&]() { if (counts) cudaFreeAsync(counts, st); if (offsets) cudaFreeAsync(offsets, st); if (idx_all) cudaFreeAsync(idx_all, st); if (dtotal) cudaFreeAsync(dtotal, st); };
  if (n > 0 && cap > 0) {
    const uint64_t nseg = (n + SEG - 1) / SEG;
    if (cudaMallocAsync((void**)&counts, nseg * 8, st) != cudaSuccess || cudaMallocAsync((void**)&offsets, nseg * 8, st) != cudaSuccess ||
        cudaMallocAsync((void**)&dtotal, 8, st) != cudaSuccess) { cudaGetLastError(); cleanup(); return fail(RM_OOM, "find: scratch allocation failed"); }
    const unsigned wblocks = (unsigned)((nseg * 32 + 255) / 256);
    find_count_kernel<T><<<wblocks, 256, 0, st>>>((const T*)src, n, counts, nseg);
    scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, nseg, dtotal);
    cudaMemcpyAsync(&found, dtotal, 8, cudaMemcpyDeviceToHost, st);
    p->host_syncs.fetch_add(1, std::memory_order_relaxed);
    cudaStreamSynchronize(st);  // the output size is data-dependent: one 8-byte read
    count_launch(p, 2);
    if (found > 0) {
      if (cudaMallocAsync((void**)&idx_all, found * 8, st) != cudaSuccess) { cudaGetLastError(); cleanup(); return fail(RM_OOM, "find: scratch allocation failed"); }
      find_compact_kernel<T><<<wblocks, 256, 0, st>>>((const T*)src, n, offsets, nseg, idx_all);
      count_launch(p);
    }
  }
  const uint64_t cnt = std::min(cap, found);
  uint64_t oshape[2] = {cnt, 1};
  void *pl = nullptr, *pr = nullptr, *pc = nullptr, *pv = nullptr;
  rm_status s1 = alloc_tensor(p, oshape, 2, linear, &pl);
  rm_status s2 = s1 == RM_OK ? alloc_tensor(p, oshape, 2, rows, &pr) : s1;
  rm_status s3 = s2 == RM_OK ? alloc_tensor(p, oshape, 2, cols, &pc) : s2;
  rm_status s4 = s3 == RM_OK ? alloc_tensor(p, oshape, 2, values, &pv) : s3;
  if (s4 != RM_OK) {
    std::string m = last_error();
    if (s1 == RM_OK) rm_free(p, linear);
    if (s2 == RM_OK) rm_free(p, rows);
    if (s3 == RM_OK) rm_free(p, cols);
    cleanup();
    set_error("%s", m.c_str());
    return s4;
  }
  if (cnt > 0) {
    find_emit_kernel<T><<<grid1d(p, cnt), 256, 0, st>>>((const T*)src, idx_all, found, cnt, last, row_extent, (T*)pl, (T*)pr, (T*)pc, (T*)pv);
    count_launch(p);
  }
  cleanup();
  RM_LAUNCH_CHECK();
  return RM_OK;
}

}  // namespace
}  // namespace rm

using namespace rm;

RM_EXPORT rm_status rm_find(rm_provider* p, const rm_handle* a, int has_limit, uint64_t limit, int direction_last, rm_handle* linear, rm_handle* rows,
                            rm_handle* cols, rm_handle* values) {
  RM_REQUIRE(p && a && linear && rows && cols && values, RM_INVALID_ARG, "find: bad arguments");
  DeviceGuard g(p->ordinal);
  return p->precision == RM_F64 ? find_impl<double>(p, a, has_limit, limit, direction_last, linear, rows, cols, values)
                                : find_impl<float>(p, a, has_limit, limit, direction_last, linear, rows, cols, values);
}

RM_EXPORT rm_status rm_scatter_column(rm_provider* p, const rm_handle* matrix, uint64_t col_index, const rm_handle* values, rm_handle* out) {
  RM_REQUIRE(p && matrix && values && out, RM_INVALID_ARG, "scatter_column: bad arguments");
  RM_REQUIRE(matrix->rank == 2, RM_ERROR, "scatter_column: only 2D tensors supported");
  DeviceGuard g(p->ordinal);
  void *pm, *pv;
  uint64_t nv;
  RM_TRY(resolve(p, matrix, &pm, nullptr));
  RM_TRY(resolve(p, values, &pv, &nv));
  const uint64_t rows = matrix->shape[0], cols = matrix->shape[1];
  RM_REQUIRE(col_index < cols, RM_ERROR, "scatter_column: column index %llu out of bounds (%llu columns)", (unsigned long long)col_index, (unsigned long long)cols);
  RM_REQUIRE(nv == rows, RM_ERROR, "scatter_column: values length %llu does not match %llu rows", (unsigned long long)nv, (unsigned long long)rows);
  void* po;
  RM_TRY(alloc_tensor(p, matrix->shape, 2, out, &po));
  const size_t es = p->elem_size();
  if (rows * cols) RM_CUDA(cudaMemcpyAsync(po, pm, rows * cols * es, cudaMemcpyDeviceToDevice, p->stream));
  if (rows) RM_CUDA(cudaMemcpyAsync((char*)po + col_index * rows * es, pv, rows * es, cudaMemcpyDeviceToDevice, p->stream));
  return RM_OK;
}

RM_EXPORT rm_status rm_scatter_row(rm_provider* p, const rm_handle* matrix, uint64_t row_index, const rm_handle* values, rm_handle* out) {
  RM_REQUIRE(p && matrix && values && out, RM_INVALID_ARG, "scatter_row: bad arguments");
  RM_REQUIRE(matrix->rank == 2, RM_ERROR, "scatter_row: only 2D tensors supported");
  DeviceGuard g(p->ordinal);
  void *pm, *pv;
  uint64_t nv;
  RM_TRY(resolve(p, matrix, &pm, nullptr));
  RM_TRY(resolve(p, values, &pv, &nv));
  const uint64_t rows = matrix->shape[0], cols = matrix->shape[1];
  RM_REQUIRE(row_index < rows, RM_ERROR, "scatter_row: row index %llu out of bounds (%llu rows)", (unsigned long long)row_index, (unsigned long long)rows);
  RM_REQUIRE(nv == cols, RM_ERROR, "scatter_row: values length %llu does not match %llu columns", (unsigned long long)nv, (unsigned long long)cols);
  void* po;
  RM_TRY(alloc_tensor(p, matrix->shape, 2, out, &po));
  if (rows * cols) RM_CUDA(cudaMemcpyAsync(po, pm, rows * cols * p->elem_size(), cudaMemcpyDeviceToDevice, p->stream));
  if (cols) {
    if (p->precision == RM_F64) scatter_row_kernel<double><<<grid1d(p, cols), 256, 0, p->stream>>>((double*)po, rows, cols, row_index, (const double*)pv);
    else scatter_row_kernel<float><<<grid1d(p, cols), 256, 0, p->stream>>>((float*)po, rows, cols, row_index, (const float*)pv);
    RM_LAUNCH_CHECK();
    count_launch(p);
  }
  return RM_OK;
}

static rm_status index_error(const char* what, int word) {
  const int code = word >> 8, dim = word & 0xff;
  if (code == 1) return dim ? fail(RM_ERROR, "%s: subscript in dimension %d must be finite", what, dim) : fail(RM_ERROR, "%s: index must be finite", what);
  if (code == 2) return dim ? fail(RM_ERROR, "%s: subscript in dimension %d must be an integer", what, dim) : fail(RM_ERROR, "%s: index must be an integer", what);
  return dim ? fail(RM_ERROR, "%s: subscript exceeds dimension %d", what, dim) : fail(RM_ERROR, "%s: index exceeds the number of elements", what);
}

RM_EXPORT rm_status rm_sub2ind(rm_provider* p, const uint64_t* dims, const uint64_t* strides, uint32_t ndims, const rm_handle* inputs,
                               const uint8_t* scalar_mask, uint64_t len, const uint64_t* output_shape, uint32_t rank, rm_handle* out) {
  RM_REQUIRE(p && dims && strides && inputs && scalar_mask && out, RM_INVALID_ARG, "sub2ind: bad arguments");
  RM_REQUIRE(ndims >= 1 && ndims <= RM_MAX_RANK, RM_ERROR, "sub2ind: expected between 1 and %d dimensions", RM_MAX_RANK);
  RM_REQUIRE(shape_elems(output_shape, rank) == len, RM_ERROR, "sub2ind: output shape does not match subscript sizes");
  DeviceGuard g(p->ordinal);
  SubParams sp{};
  sp.ndims = ndims;
  for (uint32_t d = 0; d < ndims; ++d) {
    void* ptr;
    uint64_t n;
    RM_TRY(resolve(p, &inputs[d], &ptr, &n));
    RM_REQUIRE(scalar_mask[d] ? n >= 1 : n == len, RM_ERROR, "sub2ind: subscript %u has %llu elements, expected %llu", d + 1, (unsigned long long)n, (unsigned long long)len);
    sp.dims[d] = dims[d];
    sp.strides[d] = strides[d];
    sp.ptr[d] = ptr;
    sp.scalar[d] = scalar_mask[d] ? 1 : 0;
  }
  void* po;
  RM_TRY(alloc_tensor(p, output_shape, rank, out, &po));
  if (len == 0) return RM_OK;
  int* derr = nullptr;
  RM_CUDA(cudaMallocAsync((void**)&derr, 4, p->stream));
  RM_CUDA(cudaMemsetAsync(derr, 0, 4, p->stream));
  if (p->precision == RM_F64) sub2ind_kernel<double><<<grid1d(p, len), 256, 0, p->stream>>>(sp, len, (double*)po, derr);
  else sub2ind_kernel<float><<<grid1d(p, len), 256, 0, p->stream>>>(sp, len, (float*)po, derr);
  count_launch(p);
  int herr = 0;
  cudaMemcpyAsync(&herr, derr, 4, cudaMemcpyDeviceToHost, p->stream);
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  cudaStreamSynchronize(p->stream);  // the host raises on a bad subscript, so the call must know before returning
  cudaFreeAsync(derr, p->stream);
  RM_LAUNCH_CHECK();
  if (herr) { rm_free(p, out); return index_error("sub2ind", herr); }
  return RM_OK;
}

RM_EXPORT rm_status rm_ind2sub(rm_provider* p, const uint64_t* dims, const uint64_t* strides, uint32_t ndims, const rm_handle* indices, uint64_t total,
                               uint64_t len, const uint64_t* output_shape, uint32_t rank, rm_handle* outs) {
  RM_REQUIRE(p && dims && strides && indices && outs, RM_INVALID_ARG, "ind2sub: bad arguments");
  RM_REQUIRE(ndims >= 1 && ndims <= RM_MAX_RANK, RM_ERROR, "ind2sub: expected between 1 and %d dimensions", RM_MAX_RANK);
  RM_REQUIRE(shape_elems(output_shape, rank) == len, RM_ERROR, "ind2sub: output shape does not match the number of indices");
  DeviceGuard g(p->ordinal);
  void* pi;
  uint64_t n;
  RM_TRY(resolve(p, indices, &pi, &n));
  RM_REQUIRE(n == len, RM_ERROR, "ind2sub: index tensor has %llu elements, expected %llu", (unsigned long long)n, (unsigned long long)len);
  IndParams ip{};
  ip.ndims = ndims;
  uint32_t made = 0;
  rm_status st = RM_OK;
  for (uint32_t d = 0; d < ndims && st == RM_OK; ++d) {
    ip.dims[d] = std::max<uint64_t>(dims[d], 1);
    ip.strides[d] = std::max<uint64_t>(strides[d], 1);
    st = alloc_tensor(p, output_shape, rank, &outs[d], &ip.out[d]);
    if (st == RM_OK) ++made;
  }
  auto drop = [&]() { for (uint32_t d = 0; d < made; ++d) rm_free(p, &outs[d]); };
  if (st != RM_OK) { std::string m = last_error(); drop(); set_error("%s", m.c_str()); return st; }
  if (len == 0) return RM_OK;
  int* derr = nullptr;
  if (cudaMallocAsync((void**)&derr, 4, p->stream) != cudaSuccess) { cudaGetLastError(); drop(); return fail(RM_OOM, "ind2sub: allocation failed"); }
  cudaMemsetAsync(derr, 0, 4, p->stream);
  if (p->precision == RM_F64) ind2sub_kernel<double><<<grid1d(p, len), 256, 0, p->stream>>>((const double*)pi, len, total, ip, derr);
  else ind2sub_kernel<float><<<grid1d(p, len), 256, 0, p->stream>>>((const float*)pi, len, total, ip, derr);
  count_launch(p);
  int herr = 0;
  cudaMemcpyAsync(&herr, derr, 4, cudaMemcpyDeviceToHost, p->stream);
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  cudaStreamSynchronize(p->stream);
  cudaFreeAsync(derr, p->stream);
  if (herr) { drop(); return index_error("ind2sub", herr); }
  RM_LAUNCH_CHECK();
  return RM_OK;
}
