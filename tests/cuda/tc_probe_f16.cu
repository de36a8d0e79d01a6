// Stand-alone probe of the 3xFP16 split on the 2-CTA tcgen05 path (bae_tc.cuh: f16_convert_tile, make_desc_f16, umma_f16):
// a cluster of two CTAs computes one 256 x 128 x 64 product of FULL FP32 operands. TMA writes the FP32 tiles exactly as the
// engine does (K-major rows with SWIZZLE_64B or MN-major [16 k][32 mn] chunks with SWIZZLE_128B_ATOM_32B); every thread
// converts units into the no-swizzle FP16 hi / lo tiles; the leader issues three kind::f16 MMAs (K = 16) per K block into two
// accumulators; the epilogue recombines them. All four operand-major combinations, power-of-two operand scales, checked
// against a float64 product of the unsplit inputs.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_probe_f16 tc_probe_f16.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../bayesianautoencoders_b200/csrc/bae_tc.cuh"

using namespace bae;
using namespace bae::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int M = 256, N = 128, K = 64, NKB = K / 16;

__global__ void __launch_bounds__(256, 1) probe_f16_kernel(const CUtensorMap* maps, int aMode, int bMode, unsigned idesc, float sA, float sB,
                                                           float* C) {
  extern __shared__ uint8_t dyn[];
  __shared__ uint64_t bar_full, bar_ready, bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t* rawA = tiles;                         // [4 k-blocks][8 KB]
  uint8_t* rawB = rawA + NKB * 8192;             // [4][4 KB]   (64 columns of B per CTA)
  uint8_t* a16 = rawB + NKB * 4096;              // [4][hi 4 KB | lo 4 KB]
  uint8_t* b16 = a16 + NKB * 8192;               // [4][hi 2 KB | lo 2 KB]
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_ready, 2);   // one arrival per CTA (only the leader's copy is used)
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 256);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&bar_full, NKB * (8192 + 4096));
      for (int kb = 0; kb < NKB; ++kb) {
        if (aMode == 0) tma_load_3d(rawA + kb * 8192, &maps[0], &bar_full, kb * 16, (int)rank * 128, 0);
        else for (int c = 0; c < 4; ++c) tma_load_3d(rawA + kb * 8192 + c * 2048, &maps[0], &bar_full, (int)rank * 128 + c * 32, kb * 16, 0);
        if (bMode == 0) tma_load_3d(rawB + kb * 4096, &maps[1], &bar_full, kb * 16, (int)rank * 64, 0);
        else for (int c = 0; c < 2; ++c) tma_load_3d(rawB + kb * 4096 + c * 2048, &maps[1], &bar_full, (int)rank * 64 + c * 32, kb * 16, 0);
      }
    }
    __syncwarp();
  }
  mbar_wait(&bar_full, 0);
  for (int kb = 0; kb < NKB; ++kb) {
    f16_convert_tile(smem_u32(rawA + kb * 8192), smem_u32(a16 + kb * 8192), smem_u32(a16 + kb * 8192 + 4096), aMode, sA, (int)threadIdx.x, 256, 256);
    f16_convert_tile(smem_u32(rawB + kb * 4096), smem_u32(b16 + kb * 4096), smem_u32(b16 + kb * 4096 + 2048), bMode, sB, (int)threadIdx.x, 256, 128);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive_leader(&bar_ready);
  if (warp == 1 && rank == 0) {
    mbar_wait_cluster(&bar_ready, 0);
    fence_after_sync();
    if (elect_one()) {
      for (int kb = 0; kb < NKB; ++kb) {
        const uint64_t ah = make_desc_f16(smem_u32(a16 + kb * 8192)), al = make_desc_f16(smem_u32(a16 + kb * 8192 + 4096));
        const uint64_t bh = make_desc_f16(smem_u32(b16 + kb * 4096)), bl = make_desc_f16(smem_u32(b16 + kb * 4096 + 2048));
        umma_f16(tmem, ah, bh, idesc, kb ? 1u : 0u);
        umma_f16(tmem + N, al, bh, idesc, kb ? 1u : 0u);
        umma_f16(tmem + N, ah, bl, idesc, 1u);
      }
      umma_commit(&bar_done);
    }
    __syncwarp();
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&bar_done, 0);
    fence_after_sync();
    const int r = (int)rank * 128 + q * 32 + lane;
    const float c0 = 1.0f / (sA * sB), c1 = c0 * (1.0f / 2048.0f);
    for (int n0 = 0; n0 < N; n0 += 32) {
      uint32_t v[32], w[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + n0, v);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + N + n0, w);
      for (int j = 0; j < 32; ++j) C[r * N + n0 + j] = fmaf(__uint_as_float(w[j]), c1, __uint_as_float(v[j]) * c0);
    }
    fence_before_sync();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<float> A(M * K), B(N * K);
  srand(7);
  // magnitudes spread over ~2^20 below the operand's bound: small elements exercise the FP16 subnormal range of hi
  auto rnd = [](float bound) { const float u = (rand() % 2000001 - 1000000) / 1000000.0f; const int sh = rand() % 21; return bound * u * ldexpf(1.0f, -sh * (rand() % 2)); };
  const float boundA = 3.7e-6f, boundB = 11.3f;    // delta-like and weight-like operands
  for (auto& x : A) x = rnd(boundA);
  for (auto& x : B) x = rnd(boundB);
  auto exp_of = [](float amax) { int e; frexpf(amax, &e); return 14 - e; };   // amax * 2^e in [2^13, 2^14)
  const float sA = ldexpf(1.0f, exp_of(boundA)), sB = ldexpf(1.0f, exp_of(boundB));
  std::vector<double> ref(M * N, 0.0), mag(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    double s = 0, a = 0;
    for (int k = 0; k < K; ++k) { s += (double)A[i * K + k] * B[j * K + k]; a += fabs((double)A[i * K + k] * B[j * K + k]); }
    ref[i * N + j] = s; mag[i * N + j] = a;
  }
  const int smem = NKB * (8192 + 4096) * 2 + 1024;
  CK(cudaFuncSetAttribute(probe_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int bad = 0;
  for (int aMode = 0; aMode < 2; ++aMode) for (int bMode = 0; bMode < 2; ++bMode) {
    std::vector<float> As(M * K), Bs(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) As[aMode ? k * M + i : i * K + k] = A[i * K + k];
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) Bs[bMode ? k * N + j : j * K + k] = B[j * K + k];
    float *dA, *dB, *dC; CUtensorMap* dM;
    CK(cudaMalloc(&dA, As.size() * 4)); CK(cudaMalloc(&dB, Bs.size() * 4)); CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dM, 2 * sizeof(CUtensorMap)));
    CK(cudaMemset(dC, 0, M * N * 4));
    CK(cudaMemcpy(dA, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap hm[2];
    for (int o = 0; o < 2; ++o) {
      const int mode = o ? bMode : aMode; const int rows = o ? N : M; float* p = o ? dB : dA;
      cuuint64_t dims[3]; cuuint64_t str[2]; cuuint32_t box[3]; cuuint32_t es[3] = {1, 1, 1};
      if (mode == 0) { dims[0] = K; dims[1] = rows; dims[2] = 1; str[0] = K * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 16; box[1] = o ? 64 : 128; box[2] = 1; }
      else { dims[0] = rows; dims[1] = K; dims[2] = 1; str[0] = rows * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 32; box[1] = 16; box[2] = 1; }
      CUresult r = enc(&hm[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       mode ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CK(cudaMemcpy(dM, hm, sizeof(hm), cudaMemcpyHostToDevice));
    // kind::f16: D = F32 (1 << 4), A and B formats F16 (0), majors, N >> 3, M >> 4
    const unsigned idesc = (1u << 4) | ((unsigned)aMode << 15) | ((unsigned)bMode << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const CUtensorMap* cm = dM;
    CK(cudaLaunchKernelEx(&cfg, probe_f16_kernel, cm, aMode, bMode, idesc, sA, sB, dC));
    CK(cudaDeviceSynchronize());
    std::vector<float> Ch(M * N);
    CK(cudaMemcpy(Ch.data(), dC, M * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst_abs = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) {
      worst = fmax(worst, fabs(Ch[i] - ref[i]) / mag[i]); worst_abs = fmax(worst_abs, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i]));
    }
    printf("3xFP16 aMode %d bMode %d: max |err| / sum|a b| = %.3e   (max |err| %.3e, max |ref| %.3e)\n", aMode, bMode, worst, worst_abs, maxref);
    if (!(worst < 4e-7)) bad++;
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dM);
  }
  printf(bad ? "PROBE_F16 FAILED (%d)\n" : "PROBE_F16 OK\n", bad);
  return bad ? 2 : 0;
}
