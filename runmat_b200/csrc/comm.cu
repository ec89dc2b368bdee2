// Multi-GPU exchange for the sharded paths: one process per GPU, one NCCL communicator per provider.
//
// The hot path has no data-path collective; what crosses GPUs is the final scalar of a sharded reduction (the per-rank
// payoff sum of the Monte-Carlo, the per-rank sum of an elementwise chain, an MSE partial). `rm_comm_allreduce_sum` issues
// that exchange on a dedicated communication stream and hands back a handle whose "ready" event the compute stream waits on
// only when the value is used (the same lazy mechanism uploads use), so the next step's kernels run under the collective.
// NCCL is bound at run time (dlopen of the libnccl the process already carries, e.g. torch's), never at link time: a
// single-GPU user of librm_accel_b200.so needs no NCCL at all.
//
// Peer-memory exchange (rm_comm_p2p_*). The scalar exchange above is latency-, not bandwidth-bound: a NCCL all-reduce of 8
// bytes is a kernel launch + a multi-hop LL protocol whose CTAs must become co-resident on every rank (measured r1: +175 us
// at N=4, +280 us at N=8 on a 123 us step). Over NVSwitch every GPU can store straight into every peer's memory, so the
// exchange becomes: PUBLISH = rank r stores its partial into slot [bank][r] of EVERY rank's slot buffer (CUDA IPC mapping of
// a cudaMalloc'ed 2 KB buffer) followed by a system-scope release store of the step number; COMBINE = each rank waits
// (acquire loads on its OWN memory, bounded) until all N flags of the bank carry the step, then folds the N values in rank
// order, so every rank gets the bit-identical sum independent of arrival order. The publish is fused into the tail of the
// producing kernel (the generated reduction's last block: rm_fused_reduction_allreduce; rm_payoff_partial_sum feeds the
// stand-alone publish kernel). The combine is LAZY and runs on the compute stream: it is enqueued when the result handle is first
// used (resolve / free: provider.cu) or, at the latest, right before publish(t + 4). Nothing crosses streams, so the compute
// stream stays a plain in-order chain of kernels (r07, N = 2: the event edges to and from a communication stream cost ~6 us per
// 119 us step and broke the programmatic-dependent-launch overlap of consecutive kernels), and a consumer that reads the sum
// several steps later never waits at all. Bank reuse: 8 banks; publish(t) is ordered after this rank's combine(t-4), and a
// peer's publish(t) therefore proves its combine(t-4) is done, so when rank r overwrites bank (t mod 8) with step t every peer
// has already folded step t-8 (r's publish(t) follows r's combine(t-4), which saw the peer's publish(t-4), which followed the
// peer's combine(t-8)). Every combine runs, even for a handle that is freed unused: it is the protocol's flow-control step.
#include <dlfcn.h>

#include "common.h"

namespace rm {
namespace {

struct NcclUniqueId { char internal[128]; };
using ncclComm_t = void*;
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("RUNMAT_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) { api.error = "libnccl.so.2 not found (set RUNMAT_B200_NCCL_LIB)"; return; }
    auto sym = [&](const char* s) { void* f = dlsym(api.lib, s); if (!f && api.error.empty()) api.error = std::string("missing NCCL symbol ") + s; return f; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}

rm_status nccl_ready() {
  NcclApi& a = nccl();
  if (!a.error.empty() || !a.lib) return fail(RM_UNSUPPORTED, "comm: %s", a.error.empty() ? "NCCL unavailable" : a.error.c_str());
  return RM_OK;
}

cudaEvent_t take_event(rm_provider* p) {
  cudaEvent_t ev = nullptr;
  {
    std::lock_guard<std::mutex> lk(p->ev_mu);
    if (!p->event_pool.empty()) { ev = p->event_pool.back(); p->event_pool.pop_back(); }
  }
  if (!ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  return ev;
}
void give_event(rm_provider* p, cudaEvent_t ev) {
  std::lock_guard<std::mutex> lk(p->ev_mu);
  p->event_pool.push_back(ev);
}

// ---- peer-memory exchange ---------------------------------------------------------------------------------------------
constexpr int P2P_MAXR = 16, P2P_BANKS = 8, P2P_LAG = 4;
struct P2PSlots {
  double vals[P2P_BANKS][P2P_MAXR];
  unsigned long long flags[P2P_BANKS][P2P_MAXR];
};
struct P2PState {
  P2PSlots* mine = nullptr;           // cudaMalloc (IPC-exportable; pool memory is not)
  P2PSlots* peers[P2P_MAXR] = {};     // peers[rank] == mine
  P2PSlots** d_peers = nullptr;       // device copy of `peers` (what the kernels index)
  int* d_err = nullptr;               // set by a combine whose bounded wait ran out
  int world = 1, rank = 0;
  bool connected = false;
  uint64_t step = 0;
  uint64_t ring_id[P2P_BANKS] = {};     // buffer that receives the combine of step (ring_step1 - 1)
  uint64_t ring_step1[P2P_BANKS] = {};  // step + 1, 0 = empty
};

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
// stand-alone publish (for partials produced by a kernel that does not carry the fused tail)
// Fold of one exchange step by ONE WARP (lanes < n wait for one rank's flag each, bounded; the N values are then folded in rank
// order through shuffles, so every rank gets the bit-identical sum). Used by the stand-alone combine kernel and, fused in front of
// a publish, by the publish kernel below; the generated reduction kernel carries the same code (fusion_lower.cpp combine_prev).
template <typename T>
__device__ __forceinline__ void p2p_fold_warp(const P2PSlots* __restrict__ mine, int n, unsigned long long step, T* __restrict__ out, int* __restrict__ err) {
  const int q = threadIdx.x & 31;
  const int b = (int)(step % P2P_BANKS);
  int bad = 0;
  double v = 0.0;
  if (q < n) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->flags[b][q]) < step + 1) {
      if (clock64() - t0 > 20000000000LL) { bad = 1; break; }  // ~10 s: a dead peer raises the error flag, never hangs the GPU
      __nanosleep(100);
    }
    v = *(volatile const double*)&mine->vals[b][q];
  }
  bad = __any_sync(0xffffffffu, bad);
  double s = 0.0;
  for (int r = 0; r < n; ++r) s += __shfl_sync(0xffffffffu, v, r);  // rank order
  if (q == 0) {
    if (bad) atomicExch(err, 1);
    *out = (T)(bad ? __longlong_as_double(0x7ff8000000000000LL) : s);  // exchanged and folded in f64 whatever the storage type
  }
}
template <typename T>
__global__ void p2p_publish_kernel(P2PSlots* const* __restrict__ peers, int n, int rank, unsigned long long step, const T* __restrict__ value,
                                   T* __restrict__ prev_dst, unsigned long long prev_step1, int* __restrict__ err) {
  const int q = threadIdx.x;
  if (prev_step1) { p2p_fold_warp<T>(peers[rank], n, prev_step1 - 1, prev_dst, err); __syncwarp(); }
  if (q >= n) return;
  const int b = (int)(step % P2P_BANKS);
  P2PSlots* dst = peers[q];
  dst->vals[b][rank] = (double)*value;
  st_release_sys(&dst->flags[b][rank], step + 1);
}
template <typename T>
__global__ void p2p_combine_kernel(const P2PSlots* __restrict__ mine, int n, unsigned long long step, T* __restrict__ out, int* __restrict__ err) {
  p2p_fold_warp<T>(mine, n, step, out, err);
}

P2PState* p2p_state(rm_provider* p) { return (P2PState*)p->p2p; }

}  // namespace

void p2p_enqueue_combine_locked(rm_provider* p, uint64_t step, void* dst) {
  P2PState* s = p2p_state(p);
  if (!s || !s->connected) return;
  if (p->precision == RM_F64) p2p_combine_kernel<double><<<1, 32, 0, p->stream>>>(s->mine, s->world, step, (double*)dst, s->d_err);
  else p2p_combine_kernel<float><<<1, 32, 0, p->stream>>>(s->mine, s->world, step, (float*)dst, s->d_err);
  cudaGetLastError();
  count_launch(p);
}

// The still-pending combine of `step` (if its handle has not been used or freed yet): enqueued as a kernel of its own, or -- with
// `take` -- handed to the caller, who fuses it in front of the publish it is about to enqueue. Caller holds p->comm_mu.
static void p2p_force(rm_provider* p, P2PState* s, uint64_t step, void** take = nullptr) {
  const int b = (int)(step % P2P_BANKS);
  if (s->ring_step1[b] != step + 1) return;
  s->ring_step1[b] = 0;
  std::lock_guard<std::mutex> lk(p->mu);
  auto it = p->buffers.find(s->ring_id[b]);
  if (it == p->buffers.end() || it->second.p2p_step1 != step + 1) return;  // already enqueued by resolve() / rm_free()
  if (take) *take = it->second.ptr;
  else p2p_enqueue_combine_locked(p, step, it->second.ptr);
  it->second.p2p_step1 = 0;
}

// Publish context handed to the generated reduction kernel (fused.cu): returns false when no peer exchange is connected.
// The caller holds p->comm_mu from this call until p2p_finish().
bool p2p_begin(rm_provider* p, P2PPublish* pub) {
  P2PState* s = p2p_state(p);
  if (!s || !s->connected) return false;
  const uint64_t t = s->step;
  // bank reuse rule: publish(t) is ordered after this rank's combine(t - LAG) -- fused into the publishing kernel itself (the last
  // block folds step t - LAG into its handle, then publishes step t): in steady state the exchange costs no kernel of its own
  // (r15, N = 2: a separate 1-CTA combine kernel per step sat between two programmatically-overlapped kernels and cost ~8 us)
  pub->prev_dst = nullptr;
  pub->prev_step1 = 0;
  if (t >= (uint64_t)P2P_LAG) {
    void* dst = nullptr;
    p2p_force(p, s, t - P2P_LAG, &dst);
    if (dst) { pub->prev_dst = dst; pub->prev_step1 = t - P2P_LAG + 1; }
  }
  pub->err = s->d_err;
  pub->peers = (void* const*)s->d_peers;
  pub->n = (uint32_t)s->world;
  pub->rank = (uint32_t)s->rank;
  pub->step = t;
  return true;
}
// After the producer (with its fused or stand-alone publish) has been enqueued on the compute stream: hand out a fresh 1x1 handle
// whose combine is still pending (enqueued at first use, at free, or before publish(t + LAG), whichever comes first).
rm_status p2p_finish(rm_provider* p, rm_handle* out) {
  P2PState* s = p2p_state(p);
  const uint64_t t = s->step++;
  uint64_t shp[2] = {1, 1};
  void* dst;
  RM_TRY(alloc_tensor(p, shp, 2, out, &dst));
  const int b = (int)(t % P2P_BANKS);
  s->ring_id[b] = out->buffer_id;
  s->ring_step1[b] = t + 1;
  std::lock_guard<std::mutex> lk(p->mu);
  auto it = p->buffers.find(out->buffer_id);
  if (it != p->buffers.end()) it->second.p2p_step1 = t + 1;
  return RM_OK;
}

namespace {

void p2p_destroy(rm_provider* p) {
  P2PState* s = p2p_state(p);
  if (!s) return;
  for (int q = 0; q < s->world; ++q)
    if (s->peers[q] && s->peers[q] != s->mine) cudaIpcCloseMemHandle(s->peers[q]);
  if (s->d_peers) cudaFree(s->d_peers);
  if (s->d_err) cudaFree(s->d_err);
  if (s->mine) cudaFree(s->mine);
  cudaGetLastError();
  delete s;
  p->p2p = nullptr;
}

rm_status ensure_comm_stream(rm_provider* p) {
  if (p->comm_stream) return RM_OK;
  // Highest priority: the exchange's few CTAs must get an SM slot as soon as any compute CTA retires.
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  RM_CUDA(cudaStreamCreateWithPriority(&p->comm_stream, cudaStreamNonBlocking, prio_greatest));
  return RM_OK;
}

}  // namespace

void comm_destroy(rm_provider* p) {
  if (p->comm_stream) cudaStreamSynchronize(p->comm_stream);
  p2p_destroy(p);
  if (p->nccl_comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)p->nccl_comm);
  p->nccl_comm = nullptr;
  if (p->comm_stream) { cudaStreamDestroy(p->comm_stream); p->comm_stream = nullptr; }
  p->comm_world = 1;
  p->comm_rank = 0;
}

}  // namespace rm

using namespace rm;

RM_EXPORT rm_status rm_comm_unique_id(uint8_t* out, uint32_t len) {
  RM_REQUIRE(out && len >= RM_COMM_ID_BYTES, RM_INVALID_ARG, "comm_unique_id: need a %d-byte buffer", RM_COMM_ID_BYTES);
  RM_TRY(nccl_ready());
  NcclUniqueId id;
  const int r = nccl().GetUniqueId(&id);
  RM_REQUIRE(r == 0, RM_ERROR, "ncclGetUniqueId: %s", nccl().GetErrorString(r));
  memcpy(out, id.internal, RM_COMM_ID_BYTES);
  return RM_OK;
}

RM_EXPORT rm_status rm_comm_init(rm_provider* p, const uint8_t* unique_id, uint32_t len, uint32_t rank, uint32_t world) {
  RM_REQUIRE(p && unique_id && len >= RM_COMM_ID_BYTES && world >= 1 && rank < world, RM_INVALID_ARG, "comm_init: bad arguments");
  RM_REQUIRE(!p->nccl_comm, RM_ERROR, "comm_init: communicator already initialised");
  RM_TRY(nccl_ready());
  DeviceGuard g(p->ordinal);
  NcclUniqueId id;
  memcpy(id.internal, unique_id, RM_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  const int r = nccl().CommInitRank(&comm, (int)world, id, (int)rank);
  RM_REQUIRE(r == 0 && comm, RM_ERROR, "ncclCommInitRank(rank %u of %u): %s", rank, world, nccl().GetErrorString(r));
  if (ensure_comm_stream(p) != RM_OK) { std::string m = last_error(); nccl().CommDestroy(comm); return fail(RM_ERROR, "comm_init: stream: %s", m.c_str()); }
  p->nccl_comm = comm;
  p->comm_rank = (int)rank;
  p->comm_world = (int)world;
  return RM_OK;
}

RM_EXPORT uint32_t rm_comm_world_size(rm_provider* p) {
  if (!p) return 1u;
  if (p->nccl_comm) return (uint32_t)p->comm_world;
  P2PState* s = p2p_state(p);
  return s && s->connected ? (uint32_t)s->world : 1u;
}

// Allocates this rank's slot buffer and returns its CUDA IPC handle (64 bytes) for the host to ship to every rank.
RM_EXPORT rm_status rm_comm_p2p_export(rm_provider* p, uint8_t* handle_out, uint32_t len) {
  RM_REQUIRE(p && handle_out && len >= RM_COMM_P2P_HANDLE_BYTES, RM_INVALID_ARG, "comm_p2p_export: need a %d-byte buffer", RM_COMM_P2P_HANDLE_BYTES);
  static_assert(sizeof(cudaIpcMemHandle_t) == RM_COMM_P2P_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(p->ordinal);
  std::lock_guard<std::mutex> lk(p->comm_mu);
  RM_REQUIRE(!p->p2p, RM_ERROR, "comm_p2p_export: already exported");
  std::unique_ptr<P2PState> s(new P2PState());
  RM_CUDA(cudaMalloc((void**)&s->mine, sizeof(P2PSlots)));
  RM_CUDA(cudaMemset(s->mine, 0, sizeof(P2PSlots)));
  RM_CUDA(cudaMalloc((void**)&s->d_peers, sizeof(P2PSlots*) * P2P_MAXR));
  RM_CUDA(cudaMalloc((void**)&s->d_err, sizeof(int)));
  RM_CUDA(cudaMemset(s->d_err, 0, sizeof(int)));
  RM_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  RM_CUDA(cudaIpcGetMemHandle(&h, s->mine));
  memcpy(handle_out, &h, sizeof h);
  p->p2p = s.release();
  return RM_OK;
}

// `all_handles` = world x 64 bytes in rank order (this rank's own entry is ignored). world == 1 connects the rank to itself
// (the full protocol on one GPU: used by the single-GPU tests).
RM_EXPORT rm_status rm_comm_p2p_connect(rm_provider* p, const uint8_t* all_handles, uint32_t len, uint32_t rank, uint32_t world) {
  RM_REQUIRE(p && all_handles && world >= 1 && world <= (uint32_t)P2P_MAXR && rank < world && len >= world * RM_COMM_P2P_HANDLE_BYTES, RM_INVALID_ARG,
             "comm_p2p_connect: bad arguments (world <= %d)", P2P_MAXR);
  DeviceGuard g(p->ordinal);
  std::lock_guard<std::mutex> lk(p->comm_mu);
  P2PState* s = p2p_state(p);
  RM_REQUIRE(s && s->mine, RM_ERROR, "comm_p2p_connect: call rm_comm_p2p_export first");
  RM_REQUIRE(!s->connected, RM_ERROR, "comm_p2p_connect: already connected");
  RM_TRY(ensure_comm_stream(p));
  for (uint32_t q = 0; q < world; ++q) {
    if (q == rank) { s->peers[q] = s->mine; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + (size_t)q * RM_COMM_P2P_HANDLE_BYTES, sizeof h);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (uint32_t k = 0; k < q; ++k) if (k != rank && s->peers[k]) { cudaIpcCloseMemHandle(s->peers[k]); s->peers[k] = nullptr; }
      return fail(RM_UNSUPPORTED, "comm_p2p_connect: cudaIpcOpenMemHandle(rank %u): %s", q, cudaGetErrorString(e));
    }
    s->peers[q] = (P2PSlots*)ptr;
  }
  RM_CUDA(cudaMemcpy(s->d_peers, s->peers, sizeof(P2PSlots*) * P2P_MAXR, cudaMemcpyHostToDevice));
  s->world = (int)world;
  s->rank = (int)rank;
  s->connected = true;
  return RM_OK;
}
RM_EXPORT int rm_comm_p2p_connected(rm_provider* p) {
  P2PState* s = p ? p2p_state(p) : nullptr;
  return s && s->connected ? 1 : 0;
}
// 1 when a combine's bounded wait ran out (a peer never published): waits for the communication stream.
RM_EXPORT rm_status rm_comm_p2p_error(rm_provider* p, int32_t* err) {
  RM_REQUIRE(p && err, RM_INVALID_ARG, "comm_p2p_error: bad arguments");
  *err = 0;
  P2PState* s = p2p_state(p);
  if (!s || !s->connected) return RM_OK;
  DeviceGuard g(p->ordinal);
  {
    std::lock_guard<std::mutex> lk(p->comm_mu);
    for (uint64_t u = s->step >= (uint64_t)P2P_BANKS ? s->step - P2P_BANKS : 0; u < s->step; ++u) p2p_force(p, s, u);
  }
  RM_CUDA(cudaStreamSynchronize(p->stream));
  p->host_syncs.fetch_add(1, std::memory_order_relaxed);
  int h = 0;
  RM_CUDA(cudaMemcpy(&h, s->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  *err = h;
  return RM_OK;
}

// out = sum over ranks of `in` (element-wise, same shape on every rank). The copy of `in` is ordered on the compute stream, the
// collective runs in place on that copy on the communication stream, and `out` becomes usable through its ready event.
RM_EXPORT rm_status rm_comm_allreduce_sum(rm_provider* p, const rm_handle* in, rm_handle* out) {
  RM_REQUIRE(p && in && out, RM_INVALID_ARG, "comm_allreduce_sum: bad arguments");
  DeviceGuard g(p->ordinal);
  void *src, *dst;
  uint64_t n;
  RM_TRY(resolve(p, in, &src, &n));
  if (n == 1) {
    // scalar exchange over peer memory (the sharded paths' final sum): publish into every peer's slot, lazy combine
    std::lock_guard<std::mutex> lk(p->comm_mu);
    P2PPublish pub;
    if (p2p_begin(p, &pub)) {
      if (p->precision == RM_F64)
        p2p_publish_kernel<double><<<1, 32, 0, p->stream>>>((P2PSlots* const*)pub.peers, (int)pub.n, (int)pub.rank, pub.step, (const double*)src, (double*)pub.prev_dst, pub.prev_step1, pub.err);
      else
        p2p_publish_kernel<float><<<1, 32, 0, p->stream>>>((P2PSlots* const*)pub.peers, (int)pub.n, (int)pub.rank, pub.step, (const float*)src, (float*)pub.prev_dst, pub.prev_step1, pub.err);
      RM_LAUNCH_CHECK();
      count_launch(p);
      return p2p_finish(p, out);
    }
  }
  RM_REQUIRE(p->nccl_comm, RM_ERROR, "comm_allreduce_sum: neither rm_comm_init nor rm_comm_p2p_connect has been called");
  RM_TRY(alloc_tensor(p, in->shape, in->rank, out, &dst));
  if (n == 0) return RM_OK;
  cudaEvent_t copied = take_event(p), done = take_event(p);
  cudaError_t e = cudaMemcpyAsync(dst, src, n * p->elem_size(), cudaMemcpyDeviceToDevice, p->stream);
  if (e == cudaSuccess) e = cudaEventRecord(copied, p->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(p->comm_stream, copied, 0);
  give_event(p, copied);  // the wait above captured this record; the event may be re-recorded
  int r = 0;
  if (e == cudaSuccess) r = nccl().AllReduce(dst, dst, (size_t)n, p->precision == RM_F64 ? kNcclDouble : kNcclFloat, kNcclSum, (ncclComm_t)p->nccl_comm, p->comm_stream);
  if (e == cudaSuccess && r == 0) e = cudaEventRecord(done, p->comm_stream);
  if (e != cudaSuccess || r != 0) {
    give_event(p, done);
    rm_free(p, out);
    return e != cudaSuccess ? fail(RM_ERROR, "comm_allreduce_sum: %s", cudaGetErrorString(e)) : fail(RM_ERROR, "ncclAllReduce: %s", nccl().GetErrorString(r));
  }
  {
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->buffers.find(out->buffer_id);
    if (it != p->buffers.end()) it->second.ready = done; else { cudaStreamWaitEvent(p->stream, done, 0); }
  }
  return RM_OK;
}

// Orders the compute stream after every collective issued so far (e.g. before closing a timed region).
RM_EXPORT rm_status rm_comm_fence(rm_provider* p) {
  RM_REQUIRE(p, RM_INVALID_ARG, "comm_fence: null provider");
  DeviceGuard g(p->ordinal);
  if (P2PState* s = p2p_state(p)) {
    // peer-memory exchange: enqueue every combine that is still pending (ascending step order) on the compute stream
    std::lock_guard<std::mutex> lk(p->comm_mu);
    if (s->connected)
      for (uint64_t u = s->step >= (uint64_t)P2P_BANKS ? s->step - P2P_BANKS : 0; u < s->step; ++u) p2p_force(p, s, u);
  }
  if (!p->comm_stream || !p->nccl_comm) return RM_OK;
  cudaEvent_t ev = take_event(p);
  RM_CUDA(cudaEventRecord(ev, p->comm_stream));
  RM_CUDA(cudaStreamWaitEvent(p->stream, ev, 0));
  give_event(p, ev);
  return RM_OK;
}
