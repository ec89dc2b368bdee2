// Dev harness, 2-CTA variant: int8 x int8 -> int32 GEMM on tcgen05 with cta_group::2 -- a CTA PAIR (cluster 2x1x1) owns a 256 x 256 tile,
// each CTA stages its 128 rows of A and its 128-column half of B (32 KB per stage instead of 48 KB), the leader issues
// tcgen05.mma.cta_group::2 (UMMA 256x256x32), TMA completions of both CTAs land on the leader's barrier, commits are multicast.
// (TMA -> smem -> tcgen05.mma.kind::i8 -> TMEM -> tcgen05.ld).
// C[M][N] (row-major int32) = A[M][K] (int8, K contiguous) * B[N][K]^T (int8, K contiguous).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo i8gemm2_test.cu -o i8gemm2_test
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BM = 128, BN = 256, BK = 128, STAGES = 6;          // BM = rows per CTA (pair tile 256 x 256), BN = pair tile columns
constexpr int A_STAGE = BM * BK, B_STAGE = (BN / 2) * BK;      // bytes per CTA: 128 rows of A + 128 columns of B
constexpr int STAGE_BYTES = A_STAGE + B_STAGE;                 // 32 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers
constexpr int TMEM_COLS = 256;
constexpr int NTHREADS = 192;  // warp0 TMA, warp1 MMA/alloc, warps 2-5 epilogue

__device__ int g_error = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 24); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  atomicExch(&g_error, 1);  // bounded spin: never hang the GPU during bring-up
  return false;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B canonical layout: ((8,n),2):((8 x 16B, SBO),(1 x 16B)); SBO = 8 rows * 128 B = 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);     // start address  [0,14)
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version 1 (sm_100) [46,48)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B [61,64)
  return d;
}
__device__ __forceinline__ void mma_i8_2cta(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// commit: arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair when the MMAs issued so far retire
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// TMA load issued by either CTA of the pair into ITS OWN shared memory; the transaction bytes are credited to the LEADER's barrier
// (same offset, CTA-rank bit of the shared::cluster address cleared: cute Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
i8gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int32_t* __restrict__ C, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B needs 1024-B alignment
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int tiles_n = N / BN;
  const int m0 = (pair / tiles_n) * (2 * BM) + (int)rank * BM;  // this CTA's 128 rows of the pair's 256
  const int n0 = (pair % tiles_n) * BN;
  const int nk = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // warp-collective TMEM allocation, the same warp in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote completion or multicast commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        if (!mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1)) break;
        if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);  // the leader's barrier collects the bytes of BOTH CTAs
        tma_load_2d_pair(smem + s * STAGE_BYTES, &tmA, &full[s], kb * BK, m0);
        tma_load_2d_pair(smem + s * STAGE_BYTES + A_STAGE, &tmB, &full[s], kb * BK, n0 + (int)rank * (BN / 2));
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // instruction descriptor: c=S32, a=b=INT8, K-major both, N = 256, M = 256 (128 rows per CTA)
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        if (!mbar_wait(&full[s], (kb / STAGES) & 1)) break;
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES), b_addr = a_addr + A_STAGE;
        const uint64_t da = make_desc_k_sw128(a_addr), db = make_desc_k_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < BK / 32; ++k)
          mma_i8_2cta(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        tc_commit_pair(&empty[s]);  // frees the stage in both CTAs when these MMAs retire
      }
      tc_commit_pair(tmem_full);
    }
  } else {
    const int q = warp & 3;
    if (mbar_wait(tmem_full, 0)) {
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        int4* dst = reinterpret_cast<int4*>(C + (size_t)row * N + n0 + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory and TMEM stay alive until the leader's last MMA has retired and both epilogues are done
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, void* ptr, uint64_t rows, uint64_t k, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {k};  // bytes, dim 1
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

static int run_case(EncodeTiled enc, int M, int N, int K, bool verify, int reps) {
  std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
  uint32_t s = 12345u + M + N + K;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)((int)((s >> 16) % 129) - 64); };
  for (auto& v : hA) v = rnd();
  for (auto& v : hB) v = rnd();
  int8_t *dA, *dB;
  int32_t* dC;
  CK(cudaMalloc(&dA, hA.size()));
  CK(cudaMalloc(&dB, hB.size()));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
  CUtensorMap tmA = make_map(enc, dA, M, K, BM), tmB = make_map(enc, dB, N, K, BN / 2);
  CK(cudaFuncSetAttribute(i8gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(2 * (M / (2 * BM)) * (N / BN));  // CTA pairs
  i8gemm_kernel<<<grid, NTHREADS, SMEM_BYTES>>>(tmA, tmB, dC, M, N, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int err = 0;
  CK(cudaMemcpyFromSymbol(&err, g_error, sizeof(int)));
  if (err) { printf("[%d %d %d] kernel reported a barrier timeout\n", M, N, K); return 1; }
  int bad = 0;
  if (verify) {
    std::vector<int32_t> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < M && bad < 10; i += (M > 512 ? 37 : 1))
      for (int j = 0; j < N && bad < 10; j += (N > 512 ? 41 : 1)) {
        int64_t ref = 0;
        for (int k = 0; k < K; ++k) ref += (int)hA[(size_t)i * K + k] * (int)hB[(size_t)j * K + k];
        if ((int64_t)hC[(size_t)i * N + j] != ref) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hC[(size_t)i * N + j], (long long)ref); ++bad; }
      }
    printf("[%d %d %d] verify: %s\n", M, N, K, bad ? "FAIL" : "ok");
  }
  if (reps > 0 && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) i8gemm_kernel<<<grid, NTHREADS, SMEM_BYTES>>>(tmA, tmB, dC, M, N, K);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("[%d %d %d] %.3f ms  %.1f TOP/s\n", M, N, K, ms, 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return bad ? 1 : 0;
}

int main() {
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  int rc = 0;
  rc |= run_case(enc, 256, 256, 128, true, 0);
  rc |= run_case(enc, 256, 256, 1024, true, 0);
  rc |= run_case(enc, 256, 512, 1024, true, 0);
  rc |= run_case(enc, 1024, 1024, 2048, true, 5);
  if (!rc) rc |= run_case(enc, 8192, 8192, 8192, true, 5);
  printf(rc ? "RESULT: FAIL\n" : "RESULT: PASS\n");
  return rc;
}
