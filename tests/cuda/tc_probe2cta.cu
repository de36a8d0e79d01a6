// Stand-alone probe of the 2-CTA tcgen05 path (cta_group::2): a cluster of two CTAs computes one 256 x 128 x 64 TF32
// product. CTA r holds rows [128 r, 128 r + 128) of A and columns [64 r, 64 r + 64) of B in its own shared memory; the
// leader issues tcgen05.mma.cta_group::2 and multicasts the commit; every CTA drains its own half of the accumulator.
// Launched as a cooperative kernel with a cluster dimension (the combination the persistent step kernel would need),
// including a grid-wide barrier.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_probe2cta tc_probe2cta.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../../bayesianautoencoders_b200/csrc/bae_tc.cuh"

namespace cg = cooperative_groups;
using namespace bae;
using namespace bae::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int M = 256, N = 128, K = 64;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W2_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra W2_DONE;\n\t"
      "bra W2_LOOP;\n\t"
      "W2_DONE:\n\t}"
      ::"r"(a), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __launch_bounds__(256, 1) probe2_kernel(const CUtensorMap* maps, unsigned idesc, float* C, int* flag) {
  extern __shared__ uint8_t dyn[];
  __shared__ uint64_t bar_full, bar_ready, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = tiles;                 // [4 k-blocks][128 rows x 64 B]
  uint8_t* sB = tiles + 4 * 8192;      // [4 k-blocks][64 rows x 64 B]
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_ready, 2);   // one arrival per CTA (only the leader's copy is used)
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc2(&tmem_slot, 256);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  cg::this_grid().sync();   // cooperative launch + cluster dimension: does a grid barrier still work?
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = 1;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&bar_full, 4 * 8192 + 4 * 4096);
      for (int kb = 0; kb < 4; ++kb) {
        tma_load_3d(sA + kb * 8192, &maps[0], &bar_full, kb * 16, (int)rank * 128, 0);
        tma_load_3d(sB + kb * 4096, &maps[1], &bar_full, kb * 16, (int)rank * 64, 0);
      }
    }
    __syncwarp();
    mbar_wait(&bar_full, 0);
    // this CTA's operands have landed: tell the leader (a cluster-scope release; the tensor core of the leader reads this
    // CTA's shared memory)
    asm volatile("fence.proxy.async;" ::: "memory");
    if (elect_one()) mbar_arrive_cluster(map_to_cta(smem_u32(&bar_ready), 0));
    __syncwarp();
  }
  if (warp == 1 && rank == 0) {
    mbar_wait_cluster(&bar_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      for (int kb = 0; kb < 4; ++kb)
        for (int kk = 0; kk < 2; ++kk) {
          const uint64_t a = make_desc(smem_u32(sA + kb * 8192) + kk * 32, false);
          const uint64_t b = make_desc(smem_u32(sB + kb * 4096) + kk * 32, false);
          umma2_tf32(tmem, a, b, idesc, (kb | kk) ? 1u : 0u);
        }
      umma2_commit_mc(&bar_done, 0b11);
    }
    __syncwarp();
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&bar_done, 0);
    fence_after_sync();
    const int r = (int)rank * 128 + q * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) C[r * N + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem, 256);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<float> A(M * K), B(N * K);
  srand(5);
  auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (auto& x : A) x = tf((rand() % 2000001 - 1000000) / 1000000.0f);
  for (auto& x : B) x = tf((rand() % 2000001 - 1000000) / 1000000.0f);
  std::vector<double> ref(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k]; ref[i * N + j] = s; }
  float *dA, *dB, *dC; int* dF; CUtensorMap* dM;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dM, 2 * sizeof(CUtensorMap)));
  CK(cudaMalloc(&dF, 4)); CK(cudaMemset(dF, 0, 4)); CK(cudaMemset(dC, 0, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap hm[2];
  for (int o = 0; o < 2; ++o) {
    const int rows = o ? N : M;
    cuuint64_t dims[3] = {K, (cuuint64_t)rows, 1}; cuuint64_t str[2] = {K * 4, (cuuint64_t)rows * K * 4};
    cuuint32_t box[3] = {16, (cuuint32_t)(o ? 64 : 128), 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&hm[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, o ? dB : dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  CK(cudaMemcpy(dM, hm, sizeof(hm), cudaMemcpyHostToDevice));
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  const int smem = 4 * 8192 + 4 * 4096 + 1024;
  CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  int nclusters = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, probe2_kernel, &cfg);
  printf("cudaOccupancyMaxActiveClusters: %s, %d clusters of 2 CTAs\n", cudaGetErrorString(oe), nclusters);
  const CUtensorMap* cm = dM;
  CK(cudaLaunchKernelEx(&cfg, probe2_kernel, cm, idesc, dC, dF));
  CK(cudaDeviceSynchronize());
  std::vector<float> Ch(M * N); int flag = 0;
  CK(cudaMemcpy(Ch.data(), dC, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&flag, dF, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  double e0 = 0, e1 = 0;
  for (int i = 0; i < 128 * N; ++i) { e0 = fmax(e0, fabs(Ch[i] - ref[i])); e1 = fmax(e1, fabs(Ch[128 * N + i] - ref[128 * N + i])); }
  printf("2-CTA 256x128x64: grid barrier passed %d; max |err| %.3e (rows 0-127: %.3e, rows 128-255: %.3e), max |ref| %.3e\n", flag, maxerr, e0, e1, maxref);
  const bool ok = flag == 1 && maxerr < 1e-5 * maxref;
  printf(ok ? "PROBE2CTA OK\n" : "PROBE2CTA FAILED\n");
  return ok ? 0 : 2;
}
