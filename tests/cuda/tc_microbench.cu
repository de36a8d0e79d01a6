// Stand-alone micro-benchmarks of the primitives the tcgen05 tile engine is built from (sm_100a):
// MMA issue / execution rate, commit and mbarrier hand-off latencies, TMA issue cost and latency, TMEM drain
// rate, hi/lo conversion throughput. The numbers size the pipeline of bae_tc.cuh.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tc_microbench tc_microbench.cu
#include <cstdio>
#include <cstdlib>
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

// out[0..]: cycle counts
// ---- 1. MMA: nmma back-to-back kind::tf32 128 x N x 8 MMAs (SS), pattern 0: one accumulator, 1: cross/cross/main
__global__ void __launch_bounds__(512, 1) k_mma(int N, int nmma, int pattern, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5;
  const uint32_t tmem = *sm.tmem_base();
  // zero the operand area (generic writes -> async proxy)
  for (int i = threadIdx.x; i < 4 * kStageBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(sm.tiles)[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(sm.tiles), sb = sa + kATileBytes;
    const uint32_t la = sa + kStageBytes, lb = la + kATileBytes;
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < nmma; ++i) {
        const uint32_t kk = (i & 1) * 32u;
        if (pattern == 0) {
          umma_tf32(tmem, make_desc(sa + kk, false), make_desc(sb + kk, false), idesc, i ? 1u : 0u);
        } else if (pattern == 2) {
          // per K block (6 MMAs): main main cross x4, the next block cross x4 main main -> one accumulator switch per block
          const int blk = i / 6, r6 = i % 6;
          const bool main_first = (blk & 1) == 0;
          const bool is_main = main_first ? (r6 < 2) : (r6 >= 4);
          if (is_main) umma_tf32(tmem, make_desc(sa + kk, false), make_desc(sb + kk, false), idesc, i > 1 ? 1u : 0u);
          else if (r6 & 1) umma_tf32(tmem + N, make_desc(la + kk, false), make_desc(sb + kk, false), idesc, i > 2 ? 1u : 0u);
          else umma_tf32(tmem + N, make_desc(sa + kk, false), make_desc(lb + kk, false), idesc, 1u);
        } else if (pattern == 3) {
          // one accumulator, operands alternate between the raw and the lo tiles
          const int r = i % 3;
          if (r == 0) umma_tf32(tmem, make_desc(la + kk, false), make_desc(sb + kk, false), idesc, i ? 1u : 0u);
          else if (r == 1) umma_tf32(tmem, make_desc(sa + kk, false), make_desc(lb + kk, false), idesc, 1u);
          else umma_tf32(tmem, make_desc(sa + kk, false), make_desc(sb + kk, false), idesc, 1u);
        } else {
          const int r = i % 3;
          if (r == 0) umma_tf32(tmem + N, make_desc(la + kk, false), make_desc(sb + kk, false), idesc, i ? 1u : 0u);
          else if (r == 1) umma_tf32(tmem + N, make_desc(sa + kk, false), make_desc(lb + kk, false), idesc, 1u);
          else umma_tf32(tmem, make_desc(sa + kk, false), make_desc(sb + kk, false), idesc, i > 2 ? 1u : 0u);
        }
      }
      umma_commit(&sm.tmem_full()[0]);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&sm.tmem_full()[0], 0);
    long long t2 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  teardown(sm);
}

// ---- 1b. MMA sequences with loop-invariant descriptors (tight issue loop): which orders keep the tensor pipe busy?
//  seq 0: one accumulator; 1: cross cross main; 2: main main cross x4 | cross x4 main main; 3: strict alternation;
//  4: 12 on one accumulator, 12 on the other; 5: like 2 but every MMA re-uses ONE operand pair (operand-fetch test)
template <int kSeq>
__global__ void __launch_bounds__(512, 1) k_mma_seq(int N, int nblk, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5;
  const uint32_t tmem = *sm.tmem_base();
  for (int i = threadIdx.x; i < 4 * kStageBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(sm.tiles)[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(sm.tiles), sb = sa + kATileBytes;
    const uint64_t ah = make_desc(sa, false), bh = make_desc(sb, false);
    const uint64_t al = ah + (kStageBytes >> 4), bl = bh + (kStageBytes >> 4);
    const uint32_t dm = tmem, dc = tmem + N;
    long long t0 = clock64();
    if (elect_one()) {
      for (int b = 0; b < nblk; ++b) {   // 12 MMAs per iteration
        if (kSeq == 0) {
#pragma unroll
          for (int i = 0; i < 12; ++i) umma_tf32(dm, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
        } else if (kSeq == 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            umma_tf32(dc, al + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
            umma_tf32(dc, ah + 2 * (i & 1), bl + 2 * (i & 1), idesc, 1u);
            umma_tf32(dm, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
          }
        } else if (kSeq == 2 || kSeq == 5) {
          const uint64_t xa = kSeq == 5 ? ah : al, xb = kSeq == 5 ? bh : bl;
          umma_tf32(dm, ah, bh, idesc, 1u); umma_tf32(dm, ah + 2, bh + 2, idesc, 1u);
          umma_tf32(dc, xa, bh, idesc, 1u); umma_tf32(dc, ah, xb, idesc, 1u);
          umma_tf32(dc, xa + 2, bh + 2, idesc, 1u); umma_tf32(dc, ah + 2, xb + 2, idesc, 1u);
          umma_tf32(dc, xa, bh, idesc, 1u); umma_tf32(dc, ah, xb, idesc, 1u);
          umma_tf32(dc, xa + 2, bh + 2, idesc, 1u); umma_tf32(dc, ah + 2, xb + 2, idesc, 1u);
          umma_tf32(dm, ah, bh, idesc, 1u); umma_tf32(dm, ah + 2, bh + 2, idesc, 1u);
        } else if (kSeq == 3) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            umma_tf32(dc, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
            umma_tf32(dm, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 12; ++i) umma_tf32((b & 1) ? dc : dm, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
        }
      }
      umma_commit(&sm.tmem_full()[0]);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&sm.tmem_full()[0], 0);
    long long t2 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  teardown(sm);
}

template <int kSeq>
static void run_seq(long long* out) {
  CK(cudaFuncSetAttribute(k_mma_seq<kSeq>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  for (int N : {128, 176, 240}) {
    k_mma_seq<kSeq><<<1, 512, kSmemBytes>>>(N, 50, out);
    CK(cudaDeviceSynchronize());
    printf("mma_seq %d N=%3d: issue %.1f cyc/mma, done %.1f cyc/mma (pipe ideal %d)\n", kSeq, N, out[0] / 600.0, out[1] / 600.0, N / 2);
  }
}

// ---- 2. mbarrier primitives and hand-off latency between two warps (ping-pong)
__global__ void __launch_bounds__(512, 1) k_mbar(int iters, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ uint64_t ping, pong, solo;
  if (threadIdx.x == 0) { mbar_init(&ping, 1); mbar_init(&pong, 1); mbar_init(&solo, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (warp == 0) {
    // (a) arrive + immediate wait on the same barrier (phase completes at once): cost of arrive + passing try_wait
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(&solo);
      mbar_wait(&solo, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    // (b) ping-pong with warp 1
    ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(&ping);
      mbar_wait(&pong, ph);
      ph ^= 1;
    }
    long long t2 = clock64();
    if (lane == 0) { out[0] = (t1 - t0) / iters; out[1] = (t2 - t1) / iters; }
  } else if (warp == 1) {
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&ping, ph);
      ph ^= 1;
      if (lane == 0) mbar_arrive(&pong);
    }
  }
  __syncthreads();
  // (c) MMA -> commit -> waiting warp latency with 1 tiny MMA (N = 16), and (d) with fence.proxy.async in the loop
  if (warp == 1) {
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(16 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(sm.tiles), sb = sa + kATileBytes;
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
        umma_tf32(*sm.tmem_base(), make_desc(sa, false), make_desc(sb, false), idesc, 0u);
        umma_commit(&sm.tmem_full()[0]);
      }
      __syncwarp();
      mbar_wait(&sm.tmem_full()[0], ph);
      ph ^= 1;
      fence_after_sync();
    }
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t2 = clock64();
    if (lane == 0) { out[2] = (t1 - t0) / iters; out[3] = (t2 - t1) / iters; }
  }
  teardown(sm);
}

// ---- 3. TMA: issue cost and load latency of one K block (A 128 x 16 + B nB x 16 floats, K-major, 64-byte rows)
__global__ void __launch_bounds__(512, 1) k_tma(const CUtensorMap* maps, int nB, int iters, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const uint32_t bytes = kATileBytes + nB * 64;
    long long issue = 0, total = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      long long t0 = clock64();
      if (elect_one()) {
        mbar_expect_tx(&sm.full()[0], bytes);
        tma_load_3d(sm.tiles, &maps[0], &sm.full()[0], (i % 8) * 16, 0, 0);
        tma_load_3d(sm.tiles + kATileBytes, &maps[1], &sm.full()[0], (i % 8) * 16, 0, 0);
      }
      __syncwarp();
      long long t1 = clock64();
      mbar_wait(&sm.full()[0], ph);
      ph ^= 1;
      long long t2 = clock64();
      issue += t1 - t0;
      total += t2 - t0;
    }
    // pipelined: 4 stages in flight
    long long t3 = clock64();
    uint32_t phs = 0;
    for (int i = 0; i < iters + 4; ++i) {
      const int st = i & 3;
      if (i >= 4) { mbar_wait(&sm.full()[st], (phs >> st) & 1); phs ^= 1u << st; }
      if (i < iters && elect_one()) {
        mbar_expect_tx(&sm.full()[st], bytes);
        tma_load_3d(sm.tiles + st * kStageBytes, &maps[0], &sm.full()[st], (i % 8) * 16, 0, 0);
        tma_load_3d(sm.tiles + st * kStageBytes + kATileBytes, &maps[1], &sm.full()[st], (i % 8) * 16, 0, 0);
      }
      __syncwarp();
    }
    long long t4 = clock64();
    if (lane == 0) { out[0] = issue / iters; out[1] = total / iters; out[2] = (t4 - t3) / iters; }
  }
  teardown(sm);
}

// ---- 4. TMEM drain: nw warps (4 or 8, warps 4..) each read `chunks` x (main + cross) 32-column chunks
__global__ void __launch_bounds__(512, 1) k_ldtm(int nw, int chunks, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5;
  const uint32_t tmem = *sm.tmem_base();
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  if (warp >= 4 && warp < 4 + nw) {
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int eg = (warp - 4) >> 2;
    for (int j = 0; j < chunks; ++j) {
      uint32_t a[32], b[32];
      const uint32_t c0 = (uint32_t)((eg + 2 * j) * 32) & 255u;
      tmem_ld32_issue(tbase + c0, a);
      tmem_ld32(tbase + 256 + c0, b);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]) + __uint_as_float(b[i]);
    }
  }
  asm volatile("bar.sync 3, 512;" ::: "memory");
  long long t1 = clock64();
  if (threadIdx.x == 128) { out[0] = t1 - t0; out[1] = (long long)acc; }
  teardown(sm);
}

// ---- 5. hi/lo conversion: `nthreads` threads (from thread 64 on) convert n4 float4 from one stage into another,
// `iters` times. mode 0: LDS+STS, 1: LDS only, 2: STS only. batch = float4 in flight per thread (4 or 8).
// with_mma: warp 1 issues 128 x 240 x 8 MMAs back to back meanwhile (shared-memory operand reads compete).
template <int kBatch>
__global__ void __launch_bounds__(512, 1) k_conv(int nthreads, int n4, int iters, int mode, int with_mma, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  for (int i = threadIdx.x; i < 4 * kStageBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(sm.tiles)[i] = make_float4(1.1f, 2.2f, 3.3f, 4.4f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int ct = (int)threadIdx.x - 64;
  long long t0 = clock64();
  if (warp == 1 && with_mma) {
    const int N = 240;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint32_t sa = smem_u32(sm.tiles + 2 * kStageBytes), sb = sa + kATileBytes;
    int n = 0;
    uint32_t ph = 0;
    while (!stop) {
      if (elect_one()) {
        for (int i = 0; i < 12; ++i) umma_tf32(*sm.tmem_base(), make_desc(sa + (i & 1) * 32, false), make_desc(sb + (i & 1) * 32, false), idesc, 1u);
        umma_commit(&sm.tmem_full()[0]);
      }
      __syncwarp();
      mbar_wait(&sm.tmem_full()[0], ph);
      ph ^= 1;
      n += 12;
    }
    long long t1 = clock64();
    if (threadIdx.x == 32) { out[1] = (t1 - t0) / (n ? n : 1); out[2] = n; }
  } else if (ct >= 0 && ct < nthreads) {
    const float4* src = reinterpret_cast<const float4*>(sm.tiles);
    float4* dst = reinterpret_cast<float4*>(sm.tiles + kStageBytes);
    float4 sink = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
      for (int base = ct; base < n4; base += kBatch * nthreads) {
        float4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = base + u * nthreads;
          if (i < n4) v[u] = (mode != 2) ? src[i] : make_float4(1.5f + it, 2.5f, 3.5f, 4.5f);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = base + u * nthreads;
          if (i < n4) {
            float4 l, h;
            split_tf32(v[u].x, h.x, l.x); split_tf32(v[u].y, h.y, l.y); split_tf32(v[u].z, h.z, l.z); split_tf32(v[u].w, h.w, l.w);
            if (mode != 1) dst[i] = l; else { sink.x += l.x; sink.y += l.y; sink.z += l.z; sink.w += l.w; }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
    }
    if (sink.x + sink.y + sink.z + sink.w == 12345.f) dst[0] = sink;
    asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory");
    long long t1 = clock64();
    if (ct == 0) { out[0] = (t1 - t0) / iters; stop = 1; }
  }
  __syncthreads();
  teardown(sm);
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  long long* out;
  CK(cudaMallocManaged(&out, 64 * sizeof(long long)));
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_mbar, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_ldtm, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_conv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_conv<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));

  run_seq<0>(out); run_seq<1>(out); run_seq<2>(out); run_seq<3>(out); run_seq<4>(out); run_seq<5>(out);
  for (int pattern = 0; pattern < 0; ++pattern)
    for (int N : {128, 240}) {
      for (int nmma : {600}) {
        k_mma<<<1, 512, kSmemBytes>>>(N, nmma, pattern, out);
        CK(cudaDeviceSynchronize());
        printf("mma N=%3d pattern=%d nmma=%3d: issue %lld cyc (%.1f/mma), done %lld cyc (%.1f/mma)\n", N, pattern, nmma, out[0],
               (double)out[0] / nmma, out[1], (double)out[1] / nmma);
      }
    }
  // all SMs busy (power / clock effects do not matter for cycle counts, but check anyway)
  k_mma<<<148, 512, kSmemBytes>>>(240, 600, 1, out);
  CK(cudaDeviceSynchronize());
  printf("mma N=240 pattern=1 nmma=600 on 148 CTAs: done %.1f/mma\n", (double)out[1] / 600);

  k_mbar<<<1, 512, kSmemBytes>>>(200, out);
  CK(cudaDeviceSynchronize());
  printf("mbar: arrive+passing wait %lld cyc; ping-pong round trip %lld cyc; mma(N=16)+commit+wait %lld cyc; fence.proxy.async %lld cyc\n",
         out[0], out[1], out[2], out[3]);

  // TMA maps over a [rows][128 floats] matrix (K-major operand, 16-float boxes, SWIZZLE_64B)
  float* mat;
  const int rows = 4096, ld = 128;
  CK(cudaMalloc(&mat, (size_t)rows * ld * 4));
  CK(cudaMemset(mat, 0, (size_t)rows * ld * 4));
  CUtensorMap hm[2];
  CUtensorMap* dm;
  CK(cudaMalloc(&dm, sizeof(hm)));
  for (int nB : {128, 240}) {
    for (int i = 0; i < 2; ++i) {
      cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, 1};
      cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows * ld * 4};
      cuuint32_t box[3] = {16, (cuuint32_t)(i == 0 ? 128 : nB), 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&hm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, mat, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CK(cudaMemcpy(dm, hm, sizeof(hm), cudaMemcpyHostToDevice));
    for (int grid : {1, 148}) {
      k_tma<<<grid, 512, kSmemBytes>>>(dm, nB, 200, out);
      CK(cudaDeviceSynchronize());
      printf("tma nB=%d grid=%d: issue %lld cyc, issue+land %lld cyc, pipelined (4 stages) %lld cyc per K block (%d bytes)\n", nB, grid,
             out[0], out[1], out[2], kATileBytes + nB * 64);
    }
  }
  for (int nw : {4, 8})
    for (int chunks : {2, 4}) {
      k_ldtm<<<1, 512, kSmemBytes>>>(nw, chunks, out);
      CK(cudaDeviceSynchronize());
      const double bytes = (double)nw * chunks * 2 * 4096;
      printf("ldtm warps=%d chunks=%d: %lld cyc for %.0f KB -> %.1f B/cyc\n", nw, chunks, out[0], bytes / 1024, bytes / out[0]);
    }
  for (int with_mma : {0, 1})
    for (int nt : {96, 192})
      for (int mode : {0})
        for (int batch : {4, 8}) {
          out[1] = out[2] = 0;
          if (batch == 4) k_conv<4><<<1, 512, kSmemBytes>>>(nt, 1472, 50, mode, with_mma, out);
          else k_conv<8><<<1, 512, kSmemBytes>>>(nt, 1472, 50, mode, with_mma, out);
          CK(cudaDeviceSynchronize());
          printf("convert threads=%3d mode=%d batch=%d mma=%d: %lld cyc per 23 KB stage; concurrent MMA (128x240x8) %lld cyc each\n", nt, mode,
                 batch, with_mma, out[0], out[1]);
        }
  printf("done\n");
  return 0;
}
