// Bring-up harness (NOT part of the product library): deterministic all-reduce of ONE f64 per rank over NVLink peer memory,
// the exchange DESIGN.md §8 item 1 wants fused into the reduction's last-block finish.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o allreduce_p2p_test allreduce_p2p_test.cu && ./allreduce_p2p_test [ngpus] [steps]
//
// Single process, one stream per GPU, cudaDeviceEnablePeerAccess all-to-all (the product version maps the same `Slots` buffer
// through cudaIpcGetMemHandle / cudaIpcOpenMemHandle, one process per GPU; the kernels do not change).
//
// Protocol, per step s (bank b = s & 1):
//   publish: rank r stores its partial into EVERY peer's slots.vals[b][r], then (system-scope release) slots.flags[b][r] = s+1.
//   combine: each rank waits until its own flags[b][q] >= s+1 for all q (bounded spin -> error flag, never a hang), then folds
//            vals[b][0..n) in rank order: every rank computes the bit-identical sum, independent of arrival order.
// Bank reuse is safe as long as combine(s) precedes publish(s+2) in each rank's stream order: a peer cannot reach publish(s+2)
// before it has passed combine(s+1), which needs this rank's publish(s+1), which follows this rank's combine(s).
// Expected cost: one NVLink store round (~2-3 us) instead of a NCCL kernel launch + LL protocol (~15-25 us measured per step).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int MAXR = 16;
struct Slots {
  double vals[2][MAXR];
  unsigned long long flags[2][MAXR];
};
struct Peers { Slots* p[MAXR]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}

// In the product this is the tail of the reduction kernel's last block: `value` is the block-reduced total.
__global__ void publish_kernel(Peers peers, int n, int rank, unsigned long long step, const double* value) {
  const int q = threadIdx.x;
  if (q >= n) return;
  const int b = (int)(step & 1);
  peers.p[q]->vals[b][rank] = *value;           // peer store over NVLink (or local when q == rank)
  st_release_sys(&peers.p[q]->flags[b][rank], step + 1);
}

__global__ void combine_kernel(const Slots* mine, int n, unsigned long long step, double* out, int* err) {
  __shared__ int bad;
  const int q = threadIdx.x;
  const int b = (int)(step & 1);
  if (q == 0) bad = 0;
  __syncthreads();
  if (q < n) {
    long long spins = 0;
    while (ld_acquire_sys(&mine->flags[b][q]) < step + 1) {
      if (++spins > (1ll << 26)) { bad = 1; break; }  // ~seconds: a dead peer sets the error flag instead of hanging the GPU
    }
  }
  __syncthreads();
  if (q == 0) {
    if (bad) { atomicExch(err, 1); return; }
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += mine->vals[b][r];  // rank order: identical on every rank
    *out = s;
  }
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  int n = argc > 1 ? atoi(argv[1]) : ndev;
  const int steps = argc > 2 ? atoi(argv[2]) : 2000;
  if (n > ndev) n = ndev;
  if (n > MAXR) n = MAXR;
  if (n < 1) { printf("no CUDA device\n"); return 0; }
  std::vector<cudaStream_t> st(n);
  std::vector<Slots*> slots(n);
  std::vector<double*> val(n), out(n);
  std::vector<int*> err(n);
  for (int d = 0; d < n; ++d) {
    CK(cudaSetDevice(d));
    for (int e = 0; e < n; ++e)
      if (e != d) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, d, e));
        if (!can) { printf("no peer access %d -> %d\n", d, e); return 0; }
        cudaError_t r = cudaDeviceEnablePeerAccess(e, 0);
        if (r != cudaSuccess && r != cudaErrorPeerAccessAlreadyEnabled) CK(r);
        cudaGetLastError();
      }
    CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
    CK(cudaMalloc(&slots[d], sizeof(Slots)));
    CK(cudaMemset(slots[d], 0, sizeof(Slots)));
    CK(cudaMalloc(&val[d], 8));
    CK(cudaMalloc(&out[d], 8));
    CK(cudaMalloc(&err[d], 4));
    CK(cudaMemset(err[d], 0, 4));
  }
  Peers peers{};
  for (int d = 0; d < n; ++d) peers.p[d] = slots[d];
  for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }

  // correctness: value(rank, step) = (rank + 1) * 0.5 + step  ->  sum = 0.25*n*(n+1) + n*step
  int fails = 0;
  for (unsigned long long s = 0; s < 64; ++s) {
    for (int d = 0; d < n; ++d) {
      CK(cudaSetDevice(d));
      const double v = (d + 1) * 0.5 + (double)s;
      CK(cudaMemcpyAsync(val[d], &v, 8, cudaMemcpyHostToDevice, st[d]));
      publish_kernel<<<1, 32, 0, st[d]>>>(peers, n, d, s, val[d]);
      combine_kernel<<<1, 32, 0, st[d]>>>(slots[d], n, s, out[d], err[d]);
    }
    for (int d = 0; d < n; ++d) {
      CK(cudaSetDevice(d));
      double got;
      int e;
      CK(cudaMemcpyAsync(&got, out[d], 8, cudaMemcpyDeviceToHost, st[d]));
      CK(cudaMemcpyAsync(&e, err[d], 4, cudaMemcpyDeviceToHost, st[d]));
      CK(cudaStreamSynchronize(st[d]));
      const double want = 0.25 * n * (n + 1) + (double)n * (double)s;
      if (e || got != want) { ++fails; printf("step %llu rank %d: got %.17g want %.17g err %d\n", s, d, got, want, e); }
    }
  }
  printf("correctness: %s (%d GPUs)\n", fails ? "FAILED" : "ok", n);

  // latency: `steps` back-to-back exchanges, device time of rank 0
  cudaEvent_t e0, e1;
  CK(cudaSetDevice(0));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st[0]));
  for (unsigned long long s = 64; s < 64 + (unsigned long long)steps; ++s)
    for (int d = 0; d < n; ++d) {
      CK(cudaSetDevice(d));
      publish_kernel<<<1, 32, 0, st[d]>>>(peers, n, d, s, val[d]);
      combine_kernel<<<1, 32, 0, st[d]>>>(slots[d], n, s, out[d], err[d]);
    }
  CK(cudaSetDevice(0));
  CK(cudaEventRecord(e1, st[0]));
  for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%d exchanges: %.2f us each (publish + combine kernels, launch-bound upper bound; fused into the producer it is one store round)\n",
         steps, ms * 1e3 / steps);
  return fails != 0;
}
