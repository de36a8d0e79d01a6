// Stand-alone probe of the tcgen05 tile primitives (bae_tc.cuh): one CTA, one 128 x 128 tile, K = 64.
// Checks every operand-major combination against a host reference. Build: see tests/cuda/README in DESIGN.md.
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
probe_kernel(const CUtensorMap* maps, int aMode, int bMode, unsigned idesc, float* C, float* dumpA, float* dumpB) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem = *sm.tmem_base;
  uint8_t* sA = sm.tiles;              // [2 k-blocks][16 KB]
  uint8_t* sB = sm.tiles + 2 * 16384;  // [2 k-blocks][16 KB]
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&sm.full[0], 4 * 16384);
    for (int kb = 0; kb < 2; ++kb) {
      if (aMode == 0) tma_load_3d(sA + kb * 16384, &maps[0], &sm.full[0], kb * 32, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sA + kb * 16384 + c * 4096, &maps[0], &sm.full[0], c * 32, kb * 32, 0);
      if (bMode == 0) tma_load_3d(sB + kb * 16384, &maps[1], &sm.full[0], kb * 32, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sB + kb * 16384 + c * 4096, &maps[1], &sm.full[0], c * 32, kb * 32, 0);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(&sm.full[0], 0);
    fence_after_sync();
    const uint32_t astep = aMode ? 1024u : 32u, bstep = bMode ? 1024u : 32u;
    for (int kb = 0; kb < 2; ++kb)
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t a = make_desc(smem_u32(sA + kb * 16384) + kk * astep, aMode != 0);
        uint64_t b = make_desc(smem_u32(sB + kb * 16384) + kk * bstep, bMode != 0);
        umma_tf32(tmem, a, b, idesc, (kb | kk) ? 1u : 0u);
      }
    umma_commit(&sm.tmem_full[0]);
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full[0], 0);
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
  // dump the raw shared-memory tiles (k-block 0) for inspection
  for (int i = threadIdx.x; i < 4096; i += 256) {
    dumpA[i] = reinterpret_cast<float*>(sA)[i];
    dumpB[i] = reinterpret_cast<float*>(sB)[i];
  }
  teardown(sm);
}


// 3xTF32 with the RAW fp32 operand standing in for `hi`: correct only if the tensor core ignores the low 13 mantissa bits.
__global__ void __launch_bounds__(256, 1)
probe3x_kernel(const CUtensorMap* maps, unsigned idesc, float* C) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem = *sm.tmem_base;
  uint8_t* s0 = sm.tiles;   // A_raw[2], A_lo[2], B_raw[2], B_lo[2], 16 KB each
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&sm.full[0], 8 * 16384);
    for (int o = 0; o < 4; ++o)
      for (int kb = 0; kb < 2; ++kb) tma_load_3d(s0 + (o * 2 + kb) * 16384, &maps[o], &sm.full[0], kb * 32, 0, 0);
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(&sm.full[0], 0);
    fence_after_sync();
    for (int kb = 0; kb < 2; ++kb)
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t ar = make_desc(smem_u32(s0 + (0 * 2 + kb) * 16384) + kk * 32, false);
        uint64_t al = make_desc(smem_u32(s0 + (1 * 2 + kb) * 16384) + kk * 32, false);
        uint64_t br = make_desc(smem_u32(s0 + (2 * 2 + kb) * 16384) + kk * 32, false);
        uint64_t bl = make_desc(smem_u32(s0 + (3 * 2 + kb) * 16384) + kk * 32, false);
        umma_tf32(tmem, al, br, idesc, (kb | kk) ? 1u : 0u);
        umma_tf32(tmem, ar, bl, idesc, 1u);
        umma_tf32(tmem, ar, br, idesc, 1u);
      }
    umma_commit(&sm.tmem_full[0]);
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full[0], 0);
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


// ---- BK = 16 staging (64-byte K rows): K-major uses SWIZZLE_64B, MN-major uses [16 K-rows][128 B] chunks (BASE32B swizzle)
__device__ __forceinline__ uint64_t make_desc16(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (2048u >> 4) : 1u;
  const uint64_t sbo = 512u >> 4;                      // K-major: 8 rows x 64 B; MN-major: 4 K-rows x 128 B
  const uint64_t layout = mn_major ? 1ull : 4ull;      // SWIZZLE_128B_BASE32B : SWIZZLE_64B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__global__ void __launch_bounds__(256, 1)
probe16_kernel(const CUtensorMap* maps, int aMode, int bMode, unsigned idesc, float* C) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem = *sm.tmem_base;
  uint8_t* sA = sm.tiles;              // [4 k-blocks of 16][8 KB]
  uint8_t* sB = sm.tiles + 4 * 8192;   // [4 k-blocks][8 KB]
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&sm.full[0], 8 * 8192);
    for (int kb = 0; kb < 4; ++kb) {
      if (aMode == 0) tma_load_3d(sA + kb * 8192, &maps[0], &sm.full[0], kb * 16, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sA + kb * 8192 + c * 2048, &maps[0], &sm.full[0], c * 32, kb * 16, 0);
      if (bMode == 0) tma_load_3d(sB + kb * 8192, &maps[1], &sm.full[0], kb * 16, 0, 0);
      else for (int c = 0; c < 4; ++c) tma_load_3d(sB + kb * 8192 + c * 2048, &maps[1], &sm.full[0], c * 32, kb * 16, 0);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(&sm.full[0], 0);
    fence_after_sync();
    const uint32_t astep = aMode ? 1024u : 32u, bstep = bMode ? 1024u : 32u;
    for (int kb = 0; kb < 4; ++kb)
      for (int kk = 0; kk < 2; ++kk) {
        uint64_t a = make_desc16(smem_u32(sA + kb * 8192) + kk * astep, aMode != 0);
        uint64_t b = make_desc16(smem_u32(sB + kb * 8192) + kk * bstep, bMode != 0);
        umma_tf32(tmem, a, b, idesc, (kb | kk) ? 1u : 0u);
      }
    umma_commit(&sm.tmem_full[0]);
  }
  if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full[0], 0);
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

static int test_bk16(EncodeTiledFn enc) {
  std::vector<float> A(M * K), B(N * K);
  srand(3);
  auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (auto& x : A) x = tf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : B) x = tf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<double> ref(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k]; ref[i * N + j] = s; }
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
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("BK16 aMode %d bMode %d: max |err| %.3e (max |ref| %.3e)\n", aMode, bMode, maxerr, maxref);
    if (maxerr > 1e-3 * maxref) bad++;
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dM);
  }
  return bad;
}

static int test_raw_as_hi(EncodeTiledFn enc) {
  std::vector<float> A(M * K), B(N * K), Al(M * K), Bl(N * K);
  srand(7);
  auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (int i = 0; i < M * K; ++i) { A[i] = (rand() % 2000001 - 1000000) / 1000000.0f * 1.2345f; Al[i] = A[i] - tf(A[i]); }
  for (int i = 0; i < N * K; ++i) { B[i] = (rand() % 2000001 - 1000000) / 1000000.0f * 0.789f; Bl[i] = B[i] - tf(B[i]); }
  std::vector<double> ref(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k]; ref[i * N + j] = s; }
  float* d[4]; float* dC; CUtensorMap* dM; CUtensorMap hm[4];
  const std::vector<float>* src[4] = {&A, &Al, &B, &Bl};
  for (int o = 0; o < 4; ++o) {
    CK(cudaMalloc(&d[o], M * K * 4)); CK(cudaMemcpy(d[o], src[o]->data(), M * K * 4, cudaMemcpyHostToDevice));
    cuuint64_t dims[3] = {K, 128, 1}; cuuint64_t str[2] = {K * 4, (cuuint64_t)128 * K * 4}; cuuint32_t box[3] = {32, 128, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&hm[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d[o], dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  }
  CK(cudaMalloc(&dC, M * N * 4)); CK(cudaMalloc(&dM, sizeof(hm))); CK(cudaMemcpy(dM, hm, sizeof(hm), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  probe3x_kernel<<<1, 256, kSmemBytes>>>(dM, idesc, dC);
  CK(cudaDeviceSynchronize());
  std::vector<float> Ch(M * N);
  CK(cudaMemcpy(Ch.data(), dC, M * N * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
  printf("raw-as-hi 3xTF32: max |err| %.3e, max |ref| %.3e, ratio %.2e  (%s)\n", maxerr, maxref, maxerr / maxref,
         maxerr / maxref < 3e-6 ? "tensor core truncates the low 13 bits: raw operand usable as hi" : "NOT truncating: explicit hi needed");
  return 0;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  auto tf = [](float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; };
  for (auto& x : A) x = tf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : B) x = tf((rand() % 2001 - 1000) / 1000.0f);
  // logical A(i,k), B(j,k); reference C = A * B^T
  std::vector<double> ref(M * N, 0.0);
  for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k]; ref[i * N + j] = s; }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  int bad = 0;
  for (int aMode = 0; aMode < 2; ++aMode) for (int bMode = 0; bMode < 2; ++bMode) {
    // storage: mode 0 -> [rows][K] (K contiguous); mode 1 -> [K][rows] (rows contiguous)
    std::vector<float> As(M * K), Bs(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) As[aMode ? k * M + i : i * K + k] = A[i * K + k];
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) Bs[bMode ? k * N + j : j * K + k] = B[j * K + k];
    float *dA, *dB, *dC, *dDa, *dDb; CUtensorMap* dM;
    CK(cudaMalloc(&dA, As.size() * 4)); CK(cudaMalloc(&dB, Bs.size() * 4)); CK(cudaMalloc(&dC, M * N * 4));
    CK(cudaMalloc(&dDa, 4096 * 4)); CK(cudaMalloc(&dDb, 4096 * 4)); CK(cudaMalloc(&dM, 2 * sizeof(CUtensorMap)));
    CK(cudaMemcpy(dA, As.data(), As.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, M * N * 4));
    CUtensorMap hm[2];
    for (int o = 0; o < 2; ++o) {
      const int mode = o ? bMode : aMode; const int rows = o ? N : M; float* p = o ? dB : dA;
      cuuint64_t dims[3]; cuuint64_t str[2]; cuuint32_t box[3]; cuuint32_t es[3] = {1, 1, 1};
      if (mode == 0) { dims[0] = K; dims[1] = rows; dims[2] = 1; str[0] = K * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 32; box[1] = 128; box[2] = 1; }
      else { dims[0] = rows; dims[1] = K; dims[2] = 1; str[0] = rows * 4; str[1] = (cuuint64_t)rows * K * 4; box[0] = 32; box[1] = 32; box[2] = 1; }
      CUresult r = enc(&hm[o], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       mode ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CK(cudaMemcpy(dM, hm, sizeof(hm), cudaMemcpyHostToDevice));
    unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)aMode << 15) | ((unsigned)bMode << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    probe_kernel<<<1, 256, kSmemBytes>>>(dM, aMode, bMode, idesc, dC, dDa, dDb);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("aMode %d bMode %d: kernel error %s\n", aMode, bMode, cudaGetErrorString(e)); return 1; }
    std::vector<float> Ch(M * N), da(4096), db(4096);
    CK(cudaMemcpy(Ch.data(), dC, M * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(da.data(), dDa, 4096 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(db.data(), dDb, 4096 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int nz = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(Ch[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); nz += Ch[i] != 0.f; }
    int nza = 0, nzb = 0; for (int i = 0; i < 4096; ++i) { nza += da[i] != 0.f; nzb += db[i] != 0.f; }
    printf("aMode %d bMode %d: max |err| %.3e (max |ref| %.3e) nonzero outputs %d; smem nonzeros A %d B %d; C[0..3] %g %g %g %g ref %g %g %g %g\n",
           aMode, bMode, maxerr, maxref, nz, nza, nzb, Ch[0], Ch[1], Ch[2], Ch[3], ref[0], ref[1], ref[2], ref[3]);
    if (aMode == 1) { printf("   smem A[0..7] %g %g %g %g %g %g %g %g   expected A(0..7, k=0): %g %g %g %g\n", da[0], da[1], da[2], da[3], da[4], da[5], da[6], da[7], A[0 * K], A[1 * K], A[2 * K], A[3 * K]); }
    if (maxerr > 1e-3 * maxref) bad++;
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dDa); cudaFree(dDb); cudaFree(dM);
  }
  test_raw_as_hi(enc);
  bad += test_bk16(enc);
  printf(bad ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK\n", bad);
  return bad ? 2 : 0;
}
