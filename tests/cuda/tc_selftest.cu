// Stand-alone probe of the tcgen05 tile primitives (bae_tc.cuh): one CTA, one 128 x 128 tile, K = 64 (4 K blocks of 16).
// Checks every operand-major combination against a host reference and proves that the tensor core TRUNCATES FP32
// containers to TF32 (raw tile usable as the hi operand). nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_selftest tc_selftest.cu
#include <cstdio>
#include <cstdlib>
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

constexpr int M = 128, N = 128, K = 64;

__global__ void __launch_bounds__(256, 1)
probe16_kernel(const CUtensorMap* maps, int aMode, int bMode, unsigned idesc, float* C) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem = *sm.tmem_base();
  uint8_t* sA = sm.tiles;              // [4 k-blocks of 16][8 KB]
  uint8_t* sB = sm.tiles + 4 * 8192;   // [4 k-blocks][8 KB]
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&sm.full()[0], 8 * 8192);
    for (int kb = 0; kb < 4; ++kb) {
      if (aMode == 0) tma_load_3d(sA + kb * 8192, &maps[0], &sm.full()[0], kb * 16, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sA + kb * 8192 + c * 2048, &maps[0], &sm.full()[0], c * 32, kb * 16, 0);
      if (bMode == 0) tma_load_3d(sB + kb * 8192, &maps[1], &sm.full()[0], kb * 16, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sB + kb * 8192 + c * 2048, &maps[1], &sm.full()[0], c * 32, kb * 16, 0);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(&sm.full()[0], 0);
    fence_after_sync();
    const uint32_t astep = aMode ? 1024u : 32u, bstep = bMode ? 1024u : 32u;
    for (int kb = 0; kb < 4; ++kb)
      for (int kk = 0; kk < 2; ++kk) {
        uint64_t a = make_desc(smem_u32(sA + kb * 8192) + kk * astep, aMode != 0);
        uint64_t b = make_desc(smem_u32(sB + kb * 8192) + kk * bstep, bMode != 0);
        umma_tf32(tmem, a, b, idesc, (kb | kk) ? 1u : 0u);
      }
    umma_commit(&sm.tmem_full()[0]);
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full()[0], 0);
    fence_after_sync();
    const int r = q * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) C[r * N + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
  }
  __syncthreads();
  teardown(sm);
}

static int test_bk16(EncodeTiledFn enc, bool raw) {
  std::vector<float> A(M * K), B(N * K);
  srand(3);
  auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  // raw = false: TF32-representable inputs (layout check). raw = true: full FP32 inputs; the result must match the product
  // of TRUNCATED operands (the tensor core ignores the low 13 mantissa bits), not of rounded ones: this is what lets
  // the engine feed the raw tile as the `hi` operand of the 3xTF32 split.
  auto rn = [](float x) { unsigned u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (auto& x : A) { x = (rand() % 2000001 - 1000000) / 1000000.0f * 1.2345f; if (!raw) x = tf(x); }
  for (auto& x : B) { x = (rand() % 2000001 - 1000000) / 1000000.0f * 0.789f; if (!raw) x = tf(x); }
  std::vector<double> ref(M * N, 0.0), ref_rn(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
    double s = 0, s2 = 0;
    for (int k = 0; k < K; ++k) { s += (double)tf(A[i * K + k]) * tf(B[j * K + k]); s2 += (double)rn(A[i * K + k]) * rn(B[j * K + k]); }
    ref[i * N + j] = s; ref_rn[i * N + j] = s2;
  }
  CK(cudaFuncSetAttribute(probe16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  int bad = 0;
  for (int aMode = 0; aMode < 2; ++aMode) for (int bMode = 0; bMode < 2; ++bMode) {
    std::vector<float> As(M * K), Bs(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) As[aMode ? k * M + i : i * K + k] = A[i * K + k];
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) Bs[bMode ? k * N + j : j * K + k] = B[j * K + k];
    float *dA, *dB, *dC; CUtensorMap* dM;
    CK(cudaMalloc(&dA, As.size() * 4)); CK(cudaMalloc(&dB, Bs.size() * 4)); CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dM, 2 * sizeof(CUtensorMap)));
    CK(cudaMemcpy(dA, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap hm[2];
    for (int o = 0; o < 2; ++o) {
      const int mode = o ? bMode : aMode; const int rows = o ? N : M; float* p = o ? dB : dA;
      cuuint64_t dims[3]; cuuint64_t str[2]; cuuint32_t box[3]; cuuint32_t es[3] = {1, 1, 1};
      if (mode == 0) { dims[0] = K; dims[1] = rows; dims[2] = 1; str[0] = K * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 16; box[1] = 128; box[2] = 1; }
      else { dims[0] = rows; dims[1] = K; dims[2] = 1; str[0] = rows * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 32; box[1] = 16; box[2] = 1; }
      CUresult r = enc(&hm[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       mode ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CK(cudaMemcpy(dM, hm, sizeof(hm), cudaMemcpyHostToDevice));
    unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)aMode << 15) | ((unsigned)bMode << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    probe16_kernel<<<1, 256, kSmemBytes>>>(dM, aMode, bMode, idesc, dC);
    CK(cudaDeviceSynchronize());
    std::vector<float> Ch(M * N);
    CK(cudaMemcpy(Ch.data(), dC, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0, maxerr_rn = 0;
    for (int i = 0; i < M * N; ++i) {
      maxerr = fmax(maxerr, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); maxerr_rn = fmax(maxerr_rn, fabs(Ch[i] - ref_rn[i]));
    }
    printf("%s aMode %d bMode %d: max |err| vs truncated operands %.3e, vs rounded operands %.3e (max |ref| %.3e)\n",
           raw ? "raw FP32 inputs" : "TF32 inputs    ", aMode, bMode, maxerr, maxerr_rn, maxref);
    if (maxerr > 1e-5 * maxref) bad++;
    if (raw && maxerr_rn < 10 * maxerr) bad++;   // indistinguishable from rounding: the truncation claim would be unproven
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dM);
  }
  return bad;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  int bad = test_bk16(enc, false) + test_bk16(enc, true);
  printf(bad ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK\n", bad);
  return bad ? 2 : 0;
}
