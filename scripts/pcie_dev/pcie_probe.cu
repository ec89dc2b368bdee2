// Bring-up probe (NOT part of the product): what the host link gives pinned H2D / D2H copies alone and together, per chunk size.
// The e2e bench step is H2D-bound (268 MB up, 134 MB down per step); this separates the link's ceiling from pipeline losses.
//   nvcc -O3 -o pcie_probe pcie_probe.cu && ./pcie_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
int main() {
  const size_t MAXB = 256u << 20;
  char *h_up, *h_dn, *d_up, *d_dn;
  CK(cudaMallocHost(&h_up, MAXB)); CK(cudaMallocHost(&h_dn, MAXB));
  CK(cudaMalloc(&d_up, MAXB)); CK(cudaMalloc(&d_dn, MAXB));
  memset(h_up, 1, MAXB); memset(h_dn, 2, MAXB);
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
  for (size_t chunk : {(size_t)4 << 20, (size_t)16 << 20, (size_t)32 << 20, (size_t)256 << 20}) {
    const int n = (int)(MAXB / chunk), reps = 5;
    float up = 0, dn = 0, both_up = 0, both_dn = 0;
    for (int w = 0; w < 2; ++w) {
      CK(cudaEventRecord(e0, s1));
      for (int r = 0; r < reps; ++r) for (int i = 0; i < n; ++i) CK(cudaMemcpyAsync(d_up + i * chunk, h_up + i * chunk, chunk, cudaMemcpyHostToDevice, s1));
      CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&up, e0, e1));
      CK(cudaEventRecord(e0, s2));
      for (int r = 0; r < reps; ++r) for (int i = 0; i < n; ++i) CK(cudaMemcpyAsync(h_dn + i * chunk, d_dn + i * chunk, chunk, cudaMemcpyDeviceToHost, s2));
      CK(cudaEventRecord(e1, s2)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&dn, e0, e1));
      CK(cudaEventRecord(e0, s1)); CK(cudaEventRecord(f0, s2));
      for (int r = 0; r < reps; ++r) for (int i = 0; i < n; ++i) {
        CK(cudaMemcpyAsync(d_up + i * chunk, h_up + i * chunk, chunk, cudaMemcpyHostToDevice, s1));
        if ((i & 1) == 0) CK(cudaMemcpyAsync(h_dn + i * chunk, d_dn + i * chunk, chunk, cudaMemcpyDeviceToHost, s2));  // half the bytes down, like the bench step
      }
      CK(cudaEventRecord(e1, s1)); CK(cudaEventRecord(f1, s2));
      CK(cudaEventSynchronize(e1)); CK(cudaEventSynchronize(f1));
      CK(cudaEventElapsedTime(&both_up, e0, e1)); CK(cudaEventElapsedTime(&both_dn, f0, f1));
    }
    const double gb = (double)MAXB * reps / 1e9;
    printf("chunk %4zu MiB: H2D alone %.1f GB/s, D2H alone %.1f GB/s; together: H2D %.1f GB/s (+ D2H of half the bytes %.1f GB/s)\n", chunk >> 20, gb / (up * 1e-3),
           gb / (dn * 1e-3), gb / (both_up * 1e-3), gb / 2 / (both_dn * 1e-3));
  }
  return 0;
}
