// Bring-up harness for lu_panel_push_kernel (runmat_b200/csrc/lu_panel_push.h): factors random m x 64 panels inside one cluster,
// checks P*A = L*U, |L| <= 1, the move list against the ipiv swap sequence, and times the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o scripts/lu_dev/panel_test scripts/lu_dev/panel_test.cu
//   gpurun -- 'timeout 120 scripts/lu_dev/panel_test'
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "../../runmat_b200/csrc/lu_panel_push.h"

using namespace rm;
using namespace rm::lupush;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

static int run_case(uint64_t m, uint64_t j0, int mode) {
  // the panel sits at (j0, j0) of an n x n matrix, n = j0 + m (exercises lda != m and the global row offsets)
  const uint64_t n = j0 + m, lda = n;
  std::vector<double> hA(n * NB, 0.0);  // only the panel's columns are stored: column e of the panel = hA[e*lda ...], rows j0..n-1 used
  uint64_t s = 0x9e3779b97f4a7c15ull + m * 31 + mode;
  auto rnd = [&] { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0; };
  for (uint64_t e = 0; e < (uint64_t)NB; ++e)
    for (uint64_t r = 0; r < m; ++r) {
      double v = rnd();
      if (mode == 1 && (r % 7 == 3)) v *= 1e-8;                          // badly scaled rows
      if (mode == 2) v = std::round(v * 4.0);                            // many exact ties (and zeros)
      hA[e * lda + j0 + r] = v;
    }
  double *dA = nullptr, *dW = nullptr, *dmm = nullptr;
  unsigned long long* dipiv = nullptr;
  int* dinfo = nullptr;
  RowMoves* dmv = nullptr;
  // device layout: the kernel addresses A + j0 + j0*lda, so hand it a base pointer shifted left by j0 columns
  CK(cudaMalloc(&dA, n * NB * 8)); CK(cudaMalloc(&dW, n * NB * 8)); CK(cudaMalloc(&dmm, 16)); CK(cudaMalloc(&dipiv, n * 8)); CK(cudaMalloc(&dinfo, 8));
  CK(cudaMalloc(&dmv, sizeof(RowMoves)));
  CK(cudaMemcpy(dA, hA.data(), n * NB * 8, cudaMemcpyHostToDevice));
  const double mm0[2] = {1.7976931348623157e308, 0.0};
  unsigned cl = 1;
  while (cl * ROWS < m) cl <<= 1;
  CK(cudaFuncSetAttribute(lu_panel_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  CK(cudaFuncSetAttribute(lu_panel_push_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl); cfg.blockDim = dim3(ROWS); cfg.dynamicSmemBytes = sizeof(Smem); cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nc = 0;
  CK(cudaOccupancyMaxActiveClusters(&nc, lu_panel_push_kernel, &cfg));
  if (nc < 1) { printf("m=%llu: cluster of %u not schedulable\n", (unsigned long long)m, cl); return 0; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> times;
  double* base = dW - j0 * lda;  // so that base + j0 + j0*lda = dW + j0
  for (int it = 0; it < 12; ++it) {
    CK(cudaMemcpy(dW, dA, n * NB * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(dmm, mm0, 16, cudaMemcpyHostToDevice));
    CK(cudaMemset(dinfo, 0, 8));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, lu_panel_push_kernel, base, lda, n, j0, dipiv, dinfo, dmm, dmv));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    times.push_back(ms * 1e3f);
  }
  std::sort(times.begin(), times.end());
  std::vector<double> hW(n * NB);
  std::vector<unsigned long long> hip(n);
  RowMoves mv;
  int info[2];
  double mm[2];
  CK(cudaMemcpy(hW.data(), dW, n * NB * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hip.data(), dipiv, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&mv, dmv, sizeof(mv), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(info, dinfo, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mm, dmm, 16, cudaMemcpyDeviceToHost));
  // permutation from the move list
  std::vector<uint64_t> perm(m);
  std::iota(perm.begin(), perm.end(), 0);
  bool ok = mv.count <= 2 * NB;
  for (uint32_t k = 0; k < mv.count && ok; ++k) {
    if (mv.dst[k] < j0 || mv.dst[k] >= n || mv.src[k] < j0 || mv.src[k] >= n) { ok = false; break; }
    perm[mv.dst[k] - j0] = mv.src[k] - j0;
  }
  // ... and from the swap sequence
  std::vector<uint64_t> perm2(m);
  std::iota(perm2.begin(), perm2.end(), 0);
  for (int c = 0; c < NB && ok; ++c) { const uint64_t p = hip[j0 + c] - j0; if (p >= m) { ok = false; break; } std::swap(perm2[c], perm2[p]); }
  bool perm_match = ok && perm == perm2;
  { std::vector<uint64_t> srt = perm; std::sort(srt.begin(), srt.end()); for (uint64_t i = 0; i < m && ok; ++i) if (srt[i] != i) ok = false; }
  double amax = 0, err = 0, lmax = 0, pmin = 1e308, pmax = 0;
  for (uint64_t e = 0; e < (uint64_t)NB; ++e) {
    const double d = std::fabs(hW[e * lda + j0 + e]);
    pmin = std::min(pmin, d); pmax = std::max(pmax, d);
  }
  if (ok)
    for (uint64_t i = 0; i < m; ++i)
      for (uint64_t e = 0; e < (uint64_t)NB; ++e) {
        // (L*U)(i,e) = sum_{t <= min(i,e)} L(i,t) U(t,e), L(i,i) = 1
        double acc = 0;
        const uint64_t tmax = std::min<uint64_t>(i, e);
        for (uint64_t t = 0; t <= tmax; ++t) {
          const double l = t == i ? 1.0 : hW[t * lda + j0 + i];
          acc += l * hW[e * lda + j0 + t];
        }
        const double ref = hA[e * lda + j0 + perm[i]];
        amax = std::max(amax, std::fabs(ref));
        err = std::max(err, std::fabs(acc - ref));
        if (e < i) lmax = std::max(lmax, std::fabs(hW[e * lda + j0 + i]));
      }
  const bool pass = ok && perm_match && info[0] == (mode == 2 ? info[0] : 0) && err <= 1e-12 * std::max(amax, 1.0) * (mode == 2 ? 64 : 1) && lmax <= 1.0 + 1e-12 &&
                    (mode == 2 || (std::fabs(mm[0] - pmin) <= 1e-15 * pmin && std::fabs(mm[1] - pmax) <= 1e-15 * pmax));
  printf("m=%5llu j0=%4llu mode=%d cluster=%2u  %s  info=%d moves=%u perm_match=%d  |PA-LU|=%.2e (|A|=%.2e)  max|L|=%.6f  pivots[%.3e,%.3e] vs [%.3e,%.3e]   time median %.1f us min %.1f us\n",
         (unsigned long long)m, (unsigned long long)j0, mode, cl, pass ? "PASS" : "FAIL", info[0], mv.count, (int)perm_match, err, amax, lmax, mm[0], mm[1], pmin, pmax,
         times[times.size() / 2], times[0]);
  cudaFree(dA); cudaFree(dW); cudaFree(dmm); cudaFree(dipiv); cudaFree(dinfo); cudaFree(dmv);
  return pass ? 0 : 2;
}

int main() {
  int bad = 0;
  const uint64_t ms[] = {64, 100, 256, 257, 512, 1000, 1024, 2048, 3000, 4096};
  for (uint64_t m : ms) bad += run_case(m, m == 4096 ? 0 : 192, 0) != 0;
  bad += run_case(4096, 0, 1) != 0;
  bad += run_case(2048, 64, 2) != 0;
  bad += run_case(300, 0, 2) != 0;
  printf(bad ? "FAILED (%d cases)\n" : "ALL PASS\n", bad);
  return bad ? 1 : 0;
}
