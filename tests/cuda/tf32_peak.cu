// Sustained dense kind::tf32 rate of the tensor pipes (sm_100a): every SM pair issues tcgen05.mma.cta_group::2
// (M = 256, N = 256, K = 8, operands resident in shared memory, FP32 accumulators in TMEM) back to back for a few
// seconds. Prints TFLOP/s and the SM clock the run held (clock64 cycles / elapsed time). This is the denominator of
// `roofline.frac_of_tf32` in bench.py: the issue-rate ceiling of the MMA kind the GEMM phases use, at the clock the power
// cap leaves when the tensor pipes are saturated. No memory traffic is involved, so a real GEMM cannot exceed it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tf32_peak tf32_peak.cu && ./tf32_peak [seconds]
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../bayesianautoencoders_b200/csrc/bae_tc.cuh"

using namespace bae;
using namespace bae::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// kMode 0: one kind of MMA on two alternating accumulators (dense TF32);  kMode 1: the 3xTF32 K-block pattern of bae_tc.cuh
// (2 main + 4 cross-term MMAs per 16 floats of K), i.e. the same instruction stream with the engine's accumulator order.
template <int kMode>
__global__ void __launch_bounds__(512, 1) k_peak(int N, long long nblk, long long* out) {
  extern __shared__ uint8_t dyn[];
  Smem sm;
  setup(sm, dyn);
  const int warp = threadIdx.x >> 5;
  const uint32_t tmem = *sm.tmem_base();
  for (int i = threadIdx.x; i < 4 * kStageBytes / 16; i += blockDim.x) reinterpret_cast<float4*>(sm.tiles)[i] = make_float4(1.f, 1.f, 1.f, 1.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 1 && cluster_rank() == 0) {
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
    const uint32_t sa = smem_u32(sm.tiles), sb = sa + kATileBytes;
    const uint64_t ah = make_desc(sa, false), bh = make_desc(sb, false);
    const uint64_t al = ah + (kStageBytes >> 4), bl = bh + (kStageBytes >> 4);
    const uint32_t dm = tmem, dc = tmem + N;
    const long long t0 = clock64();
    uint32_t par = 0;
    for (long long b = 0; b < nblk; ++b) {   // 12 MMAs per iteration
      if (elect_one()) {
        if (kMode == 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            umma_tf32(dm, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
            umma_tf32(dc, ah + 2 * (i & 1), bh + 2 * (i & 1), idesc, 1u);
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            umma_tf32(dm, ah, bh, idesc, 1u); umma_tf32(dm, ah + 2, bh + 2, idesc, 1u);
            umma_tf32(dc, al, bh, idesc, 1u); umma_tf32(dc, ah, bl, idesc, 1u);
            umma_tf32(dc, al + 2, bh + 2, idesc, 1u); umma_tf32(dc, ah + 2, bl + 2, idesc, 1u);
          }
        }
        if ((b & 63) == 63) umma_commit(&sm.tmem_full()[0]);
      }
      __syncwarp();
      if ((b & 63) == 63) {   // bound the queue of outstanding MMAs: wait for the batch (both CTAs' barriers are signalled)
        mbar_wait(&sm.tmem_full()[0], par);
        par ^= 1;
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  teardown(sm);
}

template <int kMode>
static void run(int N, double seconds, long long* out, int sms) {
  auto launch = [&](long long nblk) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / 2 * 2); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    nblk = (nblk + 63) / 64 * 64;
    void* args[] = {(void*)&N, (void*)&nblk, (void*)&out};
    CK(cudaLaunchKernelExC(&cfg, (const void*)k_peak<kMode>, args));
    return nblk;
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // calibrate on a short run, then one long run
  long long nblk = launch(20000);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); nblk = launch(200000); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  const long long want = (long long)(200000.0 * seconds * 1000.0 / ms);
  CK(cudaEventRecord(e0)); nblk = launch(want); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = (double)(sms / 2) * (double)nblk * 12.0 * 2.0 * 256.0 * N * 8.0;
  const double tf = flops / (ms * 1e-3) / 1e12;
  const double mhz = (double)out[0] / (ms * 1e-3) / 1e6;
  printf("{\"mode\": \"%s\", \"N\": %d, \"seconds\": %.2f, \"tflops\": %.1f, \"sm_mhz\": %.0f, \"cycles_per_mma\": %.1f, \"pairs\": %d}\n",
         kMode == 0 ? "dense kind::tf32, cta_group::2, M=256" : "3xTF32 K-block pattern (2 main + 4 cross MMAs per 16 k)", N, ms * 1e-3, tf, mhz,
         (double)out[0] / ((double)nblk * 12.0), sms / 2);
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  CK(cudaFuncSetAttribute(k_peak<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(k_peak<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  long long* out;
  CK(cudaMallocManaged(&out, 64));
  run<0>(256, seconds, out, prop.multiProcessorCount);
  run<1>(256, seconds, out, prop.multiProcessorCount);
  run<1>(240, seconds, out, prop.multiProcessorCount);
  return 0;
}
