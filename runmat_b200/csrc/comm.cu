// Multi-GPU exchange for the sharded paths: one process per GPU, one NCCL communicator per provider.
//
// The hot path has no data-path collective; what crosses GPUs is the final scalar of a sharded reduction (the per-rank
// payoff sum of the Monte-Carlo, the per-rank sum of an elementwise chain, an MSE partial). `rm_comm_allreduce_sum` issues
// that exchange on a dedicated communication stream and hands back a handle whose "ready" event the compute stream waits on
// only when the value is used (the same lazy mechanism uploads use), so the next step's kernels run under the collective.
// NCCL is bound at run time (dlopen of the libnccl the process already carries, e.g. torch's), never at link time: a
// single-GPU user of librm_accel_b200.so needs no NCCL at all.
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

}  // namespace

void comm_destroy(rm_provider* p) {
  if (p->comm_stream) cudaStreamSynchronize(p->comm_stream);
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
  // Highest priority: the collective's few CTAs must get SM slots as soon as any compute CTA retires. At default priority the
  // all-reduce queued behind a resident 592-CTA reduction / 8192-CTA elementwise kernel only starts ~a kernel later, which eats
  // the one-step slack of a pipelined caller (r04: +25 us/step at N >= 2).
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  cudaError_t e = cudaStreamCreateWithPriority(&p->comm_stream, cudaStreamNonBlocking, prio_greatest);
  if (e != cudaSuccess) { nccl().CommDestroy(comm); return fail(RM_ERROR, "comm_init: stream: %s", cudaGetErrorString(e)); }
  p->nccl_comm = comm;
  p->comm_rank = (int)rank;
  p->comm_world = (int)world;
  return RM_OK;
}

RM_EXPORT uint32_t rm_comm_world_size(rm_provider* p) { return p && p->nccl_comm ? (uint32_t)p->comm_world : 1u; }

// out = sum over ranks of `in` (element-wise, same shape on every rank). The copy of `in` is ordered on the compute stream, the
// collective runs in place on that copy on the communication stream, and `out` becomes usable through its ready event.
RM_EXPORT rm_status rm_comm_allreduce_sum(rm_provider* p, const rm_handle* in, rm_handle* out) {
  RM_REQUIRE(p && in && out, RM_INVALID_ARG, "comm_allreduce_sum: bad arguments");
  RM_REQUIRE(p->nccl_comm, RM_ERROR, "comm_allreduce_sum: rm_comm_init has not been called");
  DeviceGuard g(p->ordinal);
  void *src, *dst;
  uint64_t n;
  RM_TRY(resolve(p, in, &src, &n));
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
  if (!p->comm_stream) return RM_OK;
  DeviceGuard g(p->ordinal);
  cudaEvent_t ev = take_event(p);
  RM_CUDA(cudaEventRecord(ev, p->comm_stream));
  RM_CUDA(cudaStreamWaitEvent(p->stream, ev, 0));
  give_event(p, ev);
  return RM_OK;
}
