// Persistent PT-Langevin step kernel for sm_100a: the whole iteration of MAIN:432-592 for every
// chain on the GPU runs as ONE cooperative launch that walks a phase program (forward GEMMs with
// fused bias/sigmoid/loss epilogues, train-mode BatchNorm, backward GEMMs with fused sigmoid-derivative
// and bias-gradient epilogues, Adam drift + Philox proposal noise, likelihood/prior/reverse-density
// reduction and the Metropolis-Hastings test), separated by grid barriers.
//
// This file holds the FP32 CUDA-core path (tiled SGEMM; 128x128 or 64x64 tiles, 256 threads);
// the tcgen05 3xTF32 tile engine for wide layers lives in bae_tc.cuh and plugs into the same phases.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "bae_common.cuh"
#include "bae_tc.cuh"

namespace cg = cooperative_groups;

namespace bae {

constexpr int BK = 16;
// The phase helpers below run on 256 threads of a CTA: all of them in the FP32 kernels, the two epilogue warpgroups
// (threads 128..383, the ones holding the large register budget) in the tcgen05 kernels. They index themselves with
// TID (a permutation of 0..255 in both cases) and synchronise on a named barrier instead of __syncthreads().
#define BLOCK_SYNC() asm volatile("bar.sync 2, 256;" ::: "memory")
#define TID (threadIdx.x & 255)
constexpr float kBnEps = 1e-5f;
constexpr float kBnMom = 0.1f;
constexpr float kB1 = 0.9f, kB2 = 0.999f, kAdamEps = 1e-8f;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (all 256 threads call). Result valid in every thread.
__device__ __forceinline__ double block_sum_d(double v, double* red /* >= 9 doubles */) {
  v = warp_sum_d(v);
  int w = TID >> 5, l = TID & 31;
  BLOCK_SYNC();
  if (l == 0) red[w] = v;
  BLOCK_SYNC();
  if (TID == 0) {
    double s = 0.0;
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    red[8] = s;
  }
  BLOCK_SYNC();
  return red[8];
}

__device__ __forceinline__ void box_muller(float u1, float u2, float& n0, float& n1) {
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// Four N(0,1) draws for (chain, step, tensor seg, quad index).
__device__ __forceinline__ void noise_quad(const DevState* S, int chain_global, int step, int seg, unsigned quad,
                                           float (&n)[4]) {
  u4 c;
  c.x = quad;
  c.y = (unsigned)seg | (STREAM_WEIGHT_NOISE << 8);
  c.z = (unsigned)step;
  c.w = (unsigned)chain_global;
  u4 r = philox4x32_10(c, S->seed_lo, S->seed_hi);
  box_muller(u01(r.x), u01(r.y), n[0], n[1]);
  box_muller(u01(r.z), u01(r.w), n[2], n[3]);
}

// Tensor table of the flat parameter vector in shared memory (filled once per kernel): the tcgen05 kernels run with the
// shared-memory carve-out at its maximum, i.e. without an L1 to cache the table in global memory.
struct SegLite { int pad_off, span, cols, ld, trainable, ksplit; float inv_ld; };   // span = rows * ld
__shared__ SegLite sh_seg[kNumSegs];

__device__ __forceinline__ void load_seg_table(const DevState* S) {
  for (int i = threadIdx.x; i < kNumSegs; i += blockDim.x) {
    const Seg& g = S->seg[i];
    sh_seg[i] = SegLite{g.pad_off, g.rows * g.ld, g.cols, g.ld, g.trainable, S->seg_ksplit[i], 1.0f / (float)g.ld};
  }
  __syncthreads();
}

__device__ __forceinline__ int find_seg(const DevState* S, int p) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < kNumSegs; ++i) s += (p >= sh_seg[i].pad_off) ? 1 : 0;
  return s;
}

// ------------------------------------------------------------------------------------------------
// FP32 tiled GEMM item: C[M x N] = A(M x K) * B(N x K)^T with a fused epilogue
// ------------------------------------------------------------------------------------------------
template <int BT>
__device__ __forceinline__ void tile_load(const float* P, int ld, int mode, int extent, int K, int t0,
                                          int k0, float4 (&r)[BT * 4 / kThreads]) {
  constexpr int NV = BT * 4 / kThreads;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int f = TID + i * kThreads;
    const float* addr;
    bool valid;
    if (mode == 0) {  // K contiguous
      int row = f % BT, k4 = f / BT;
      int gr = t0 + row, gk = k0 + k4 * 4;
      valid = (gr < extent) && (gk < K);
      addr = P + (size_t)gr * ld + gk;
    } else {  // tile dimension contiguous
      int kk = f / (BT / 4), t4 = f % (BT / 4);
      int gk = k0 + kk, gt = t0 + t4 * 4;
      valid = (gk < K) && (gt < extent);
      addr = P + (size_t)gk * ld + gt;
    }
    // plain (coherent) loads: operands are rewritten by earlier phases of the same launch, so ld.global.nc is illegal
    r[i] = valid ? *reinterpret_cast<const float4*>(addr) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int BT>
__device__ __forceinline__ void tile_store(float* __restrict__ sm /*[BK][BT]*/, int mode,
                                           const float4 (&r)[BT * 4 / kThreads]) {
  constexpr int NV = BT * 4 / kThreads;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int f = TID + i * kThreads;
    if (mode == 0) {
      int row = f % BT, k4 = f / BT;
      sm[(k4 * 4 + 0) * BT + row] = r[i].x;
      sm[(k4 * 4 + 1) * BT + row] = r[i].y;
      sm[(k4 * 4 + 2) * BT + row] = r[i].z;
      sm[(k4 * 4 + 3) * BT + row] = r[i].w;
    } else {
      int kk = f / (BT / 4), t4 = f % (BT / 4);
      *reinterpret_cast<float4*>(&sm[kk * BT + t4 * 4]) = r[i];
    }
  }
}

template <int BM, int BN>
__device__ void gemm_item(const DevState* S, const GemmDesc& g, int chain, int slot, int tile_m, int tile_n, int split, float* smem,
                          double* red) {
  constexpr int TM = BM / 16, TN = BN / 16;  // 8x8 (128) or 4x4 (64) per thread
  constexpr int GM = TM / 4, GN = TN / 4;
  // descriptor fields used inside the K loop live in locals: BLOCK_SYNC is an asm volatile with a memory clobber, after
  // which the compiler would re-load them from *g on every K tile
  const int lda = g.lda, ldb = g.ldb, aMode = g.aMode, bMode = g.bMode, gM = g.M, gN = g.N, gK = g.K, ksplit = g.ksplit;
  float* As = smem;                 // [2][BK][BM]
  float* Bs = smem + 2 * BK * BM;   // [2][BK][BN]
  const float* A = g.A + (size_t)(g.aAct ? slot : chain) * g.sA;   // activations live in the chain's slot of the resident group
  const float* B = g.B + (size_t)(g.bAct ? slot : chain) * g.sB;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int tx = TID & 15, ty = TID >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float colsum = 0.f;
  const bool do_colsum = (g.epi == EPI_GRADW) && (g.gbias != nullptr) && (tile_n == 0) && (TID < BM);

  float4 ra[BM * 4 / kThreads], rb[BN * 4 / kThreads];
  const int nk_all = (gK + BK - 1) / BK;
  const int kt0 = (int)(((long long)nk_all * split) / ksplit);          // this item's K-tile range
  const int nk = (int)(((long long)nk_all * (split + 1)) / ksplit) - kt0;
  A += (size_t)kt0 * BK * (aMode ? lda : 1);
  B += (size_t)kt0 * BK * (bMode ? ldb : 1);
  const int Krem = gK - kt0 * BK;   // loads are bounded against the remaining K
  tile_load<BM>(A, lda, aMode, gM, Krem, m0, 0, ra);
  tile_load<BN>(B, ldb, bMode, gN, Krem, n0, 0, rb);
  BLOCK_SYNC();  // previous item's smem reads are done
  tile_store<BM>(As, aMode, ra);
  tile_store<BN>(Bs, bMode, rb);
  BLOCK_SYNC();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      tile_load<BM>(A, lda, aMode, gM, Krem, m0, (kt + 1) * BK, ra);
      tile_load<BN>(B, ldb, bMode, gN, Krem, n0, (kt + 1) * BK, rb);
    }
    const float* as = As + cur * BK * BM;
    const float* bs = Bs + cur * BK * BN;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int gi = 0; gi < GM; ++gi) {
        float4 t = *reinterpret_cast<const float4*>(&as[k * BM + gi * 64 + ty * 4]);
        a[gi * 4 + 0] = t.x; a[gi * 4 + 1] = t.y; a[gi * 4 + 2] = t.z; a[gi * 4 + 3] = t.w;
      }
#pragma unroll
      for (int hi = 0; hi < GN; ++hi) {
        float4 t = *reinterpret_cast<const float4*>(&bs[k * BN + hi * 64 + tx * 4]);
        b[hi * 4 + 0] = t.x; b[hi * 4 + 1] = t.y; b[hi * 4 + 2] = t.z; b[hi * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_colsum) {
#pragma unroll
      for (int k = 0; k < BK; ++k) colsum += as[k * BM + TID];
    }
    if (kt + 1 < nk) {
      tile_store<BM>(As + (cur ^ 1) * BK * BM, aMode, ra);
      tile_store<BN>(Bs + (cur ^ 1) * BK * BN, bMode, rb);
    }
    BLOCK_SYNC();
  }

  // ---- epilogue
  const size_t split_off = (size_t)split * S->C_alloc * S->Pp;   // gradient copy of this K split (0 when ksplit == 1)
  float* C = g.C ? g.C + (size_t)(g.cAct ? slot : chain) * g.sC + split_off : nullptr;
  const float* bias = g.bias ? g.bias + (size_t)chain * g.sBias : nullptr;
  const float* aux = g.aux ? g.aux + (size_t)(g.auxAct ? slot : chain) * g.sAux : nullptr;
  float l_sse = 0.f, l_sd = 0.f;
#pragma unroll
  for (int gi = 0; gi < GM; ++gi) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int i = gi * 4 + ii;
      const int r = m0 + gi * 64 + ty * 4 + ii;
      if (r >= g.M) continue;
#pragma unroll
      for (int hi = 0; hi < GN; ++hi) {
        const int c = n0 + hi * 64 + tx * 4;
        if (c >= g.N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][hi * 4 + j];
        const int nval = min(4, g.N - c);
        float bv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias) {
          float4 t = *reinterpret_cast<const float4*>(&bias[c]);  // pads are zero
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        }
        if (aux) {
          float4 t = *reinterpret_cast<const float4*>(&aux[(size_t)r * g.ldaux + c]);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        }
        switch (g.epi) {
          case EPI_SIGMOID:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = sigmoidf_(v[j] + bv[j]);
            break;
          case EPI_BIAS:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] + bv[j];
            break;
          case EPI_LOSS:
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float o = v[j] + bv[j];
              float d = o - av[j];
              if (j < nval) { l_sse = fmaf(d, d, l_sse); l_sd += d; }
              v[j] = d * g.scale;
            }
            break;
          case EPI_DSIG:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] * av[j] * (1.0f - av[j]);
            break;
          default:
            break;
        }
        if (C) {
          float* dst = &C[(size_t)r * g.ldc + c];
          if (nval == 4) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
            for (int j = 0; j < nval; ++j) dst[j] = v[j];
          }
        }
      }
    }
  }
  if (g.epi == EPI_LOSS) {
    double s0 = block_sum_d((double)l_sse, red);
    double s1 = block_sum_d((double)l_sd, red);
    if (TID == 0) {
      double* p = g.part + ((size_t)chain * g.partStride + (size_t)tile_m * g.tn + tile_n) * 2;
      p[0] = s0;
      p[1] = s1;
    }
  }
  if (do_colsum) {
    int r = m0 + TID;
    if (r < g.M) (g.gbias + (size_t)chain * g.sGb + split_off)[r] = colsum;
  }
}

__device__ void run_gemm_phase(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, float* smem,
                               double* red) {
  const int nch = c_end - c_begin;
  const GemmDesc& ga = S->gemm[ph.g0];
  const int ta = ga.tm * ga.tn * ga.ksplit;
  const int tb = (ph.g1 >= 0) ? S->gemm[ph.g1].tm * S->gemm[ph.g1].tn * S->gemm[ph.g1].ksplit : 0;
  const int per_chain = ta + tb;
  const int total = per_chain * nch;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / per_chain;
    int t = item % per_chain;
    const GemmDesc& g = (t < ta) ? ga : S->gemm[ph.g1];
    if (t >= ta) t -= ta;
    if (g.needLang && !S->ch.is_lang[chain]) continue;
    const int split = t % g.ksplit;
    t /= g.ksplit;
    const int tile_m = t / g.tn, tile_n = t % g.tn;
    if (g.big)
      gemm_item<128, 128>(S, g, chain, chain - c_begin, tile_m, tile_n, split, smem, red);
    else
      gemm_item<64, 64>(S, g, chain, chain - c_begin, tile_m, tile_n, split, smem, red);
  }
}

// ------------------------------------------------------------------------------------------------
// Magnitude bounds behind the FP16 operand scales of the 3xFP16 split (bae_common.cuh: AmaxSlot)
// ------------------------------------------------------------------------------------------------
// CTA-wide maximum of `v` folded into amax[slot][chain] (one atomic per warp; every thread of the CTA calls this)
// (v >= 0; the lanes of a warp that are not here - columns beyond d3 in the BatchNorm phases - simply do not take part)
__device__ __forceinline__ void fold_amax(const DevState* S, int slot, int chain, float v) {
  const unsigned m = __activemask();
  const unsigned r = __reduce_max_sync(m, __float_as_uint(v));
  if ((int)(TID & 31) == __ffs(m) - 1 && r != 0u)
    atomicMax(reinterpret_cast<unsigned*>(S->amax) + (size_t)slot * S->C_alloc + chain, r);
}

// PH_WAMAX: max |w| over the six weight matrices of every chain in [c_begin, c_end) (slot AMAX_W, zeroed by the host's
// bae_amax_zero_kernel launch in front of this phase). Work item = (chain, 4096-float piece of one matrix).
__device__ void run_wamax(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, int force) {
  if (!S->f16) return;
  if (ph.arg == 1 && !force) {
    // start of an iteration inside a step program: theta changed only if the previous iteration ended with an exchange
    // round (the swap phases' own gate, seen one iteration later); the first iteration of a call is covered by the forced
    // launch the host puts in front of the program
    const int st = S->ch.step[c_begin];
    if ((st % S->swap_interval) != 0) return;
  }
  constexpr int kPiece = 4096;
  const int wsegs[6] = {SEG_D1_W, SEG_D3_W, SEG_D5_W, SEG_E0_W, SEG_E2_W, SEG_E4_W};
  int np[6], tot = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) { np[i] = (S->seg[wsegs[i]].rows * S->seg[wsegs[i]].ld + kPiece - 1) / kPiece; tot += np[i]; }
  const int total = tot * (c_end - c_begin);
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / tot;
    int p = item % tot, sg = 0;
    while (p >= np[sg]) { p -= np[sg]; ++sg; }
    const Seg& sd = S->seg[wsegs[sg]];
    const int n = sd.rows * sd.ld;
    const float* w = S->theta + (size_t)chain * S->Pp + sd.pad_off;
    float amx = 0.f;
    for (int e = p * kPiece + TID * 4; e < min(n, (p + 1) * kPiece); e += kThreads * 4) {
      const float4 v = *reinterpret_cast<const float4*>(w + e);   // rows are padded to 4 floats, pads are zero
      amx = fmaxf(amx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    fold_amax(S, AMAX_W, chain, amx);
  }
}

// Zeroes the amax slots in `mask` for chains [c_begin, c_end): launched by the host in front of the phase range that
// produces them (forward pass: Y of that pass, DOUT; backward pass: the five inner deltas; PH_WAMAX: the weights).
// `gated`: the same test as the gated PH_WAMAX phase it precedes (run_wamax).
__global__ void bae_amax_zero_kernel(const DevState* __restrict__ S, unsigned mask, int c_begin, int c_end, int gated) {
  if (gated && (S->ch.step[c_begin] % S->swap_interval) != 0) return;
  const int n = c_end - c_begin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kAmaxSlots * n; i += gridDim.x * blockDim.x) {
    const int slot = i / n, chain = c_begin + i % n;
    if ((mask >> slot) & 1u) S->amax[(size_t)slot * S->C_alloc + chain] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Pre-split operand images of the 3xFP16 split (bae_tc.cuh "operand images"): the FP16 hi / lo units the converter warps
// would produce, written once per rewrite of the source matrix in the order the tiles are consumed.
// ------------------------------------------------------------------------------------------------
// One work item: rows [o0, o0 + 8) x units [g0, g0 + 64) of a row-major FP32 matrix (unit g = columns 8 g .. 8 g + 7). Thread
// t takes row o0 + (t & 7) and units g0 + (t >> 3), g0 + 32 + (t >> 3): a warp reads 8 rows x 128 contiguous bytes and writes,
// in each image and plane, runs of 128 contiguous bytes or more (the 8 rows of a unit column).
//   imgK (contraction over the columns; columns >= ccK read as 0): unit at (g >> 1) kbsK + (o >> 3) 256 + (g & 1) 128 + (o & 7) 16
//   imgM (contraction over the rows;    columns >= ccM read as 0): unit at (o >> 4) kbsM + ((o >> 3) & 1) kbsM / 4 + g 128 + (o & 7) 16
// lo planes kbs / 2 further. Rows >= R and units beyond the columns are never written: they keep the zeros of the allocation
// (what the tensor loads' out-of-bounds fill gives the FP32 path).
__device__ __forceinline__ void split_item_to_images(const float* __restrict__ src, int ld, int R, int ccK, int ccM, float s,
                                                     unsigned char* imgK, int kbsK, unsigned char* imgM, int kbsM, int o0, int g0) {
  const int o = o0 + (int)(TID & 7);
  if (o >= R) return;
  const int ccmax = max(imgK ? ccK : 0, imgM ? ccM : 0);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int g = g0 + u * 32 + (int)(TID >> 3);
    const int c0 = g * 8;
    if (c0 >= ccmax) continue;
    const float* p = src + (size_t)o * ld + c0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v0 = *reinterpret_cast<const float4*>(p);                         // ld is a multiple of 4 and c0 < cols <= ld
    const float4 v1 = (c0 + 4 < ld) ? *reinterpret_cast<const float4*>(p + 4) : z4;
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    auto emit = [&](unsigned char* img, int cc, size_t off, int kbs) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (c0 + e < cc) ? x[e] : 0.f;
      uint4 hi, lo;
      tc::f16_split8(make_float4(y[0], y[1], y[2], y[3]), make_float4(y[4], y[5], y[6], y[7]), s, hi, lo);
      *reinterpret_cast<uint4*>(img + off) = hi;
      *reinterpret_cast<uint4*>(img + off + (size_t)(kbs >> 1)) = lo;
    };
    if (imgK && c0 < ccK) emit(imgK, ccK, (size_t)(g >> 1) * kbsK + (size_t)(o >> 3) * 256 + (size_t)(g & 1) * 128 + (size_t)(o & 7) * 16, kbsK);
    if (imgM && c0 < ccM) emit(imgM, ccM, (size_t)(o >> 4) * kbsM + (size_t)((o >> 3) & 1) * (size_t)(kbsM >> 2) + (size_t)g * 128 + (size_t)(o & 7) * 16, kbsM);
  }
}

// Images of the six weight matrices of every chain in [c_begin, c_end), from theta with the scale of the chain's CURRENT weight
// bound (amax slot AMAX_W: final when this runs - it follows the phase that folded it, PH_WAMAX or PH_ADAM1, on the stream).
// `gated`: the same test as the gated PH_WAMAX phase it follows.
__global__ void __launch_bounds__(kThreads) bae_wconv_kernel(const DevState* __restrict__ S, int c_begin, int c_end, int gated) {
  if (gated && (S->ch.step[c_begin] % S->swap_interval) != 0) return;
  const int per_chain = S->wimg_items;
  const int total = per_chain * (c_end - c_begin);
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / per_chain;
    int p = item % per_chain, w = 0;
    while (p >= S->wi[w].items) { p -= S->wi[w].items; ++w; }
    const WImg wi = S->wi[w];
    const Seg& sd = S->seg[wi.seg];
    const int ub = (((sd.cols + 7) >> 3) + 63) >> 6;   // 64-unit blocks per row
    const float s = tc::exp2i(f16_scale_exp(S->amax[(size_t)AMAX_W * S->C_alloc + chain]));
    unsigned char* slab = S->wimg + (size_t)chain * S->wimg_stride;
    split_item_to_images(S->theta + (size_t)chain * S->Pp + sd.pad_off, sd.ld, sd.rows, sd.cols, sd.cols, s, slab + wi.offK, wi.kbsK,
                         wi.offM >= 0 ? slab + wi.offM : nullptr, wi.kbsM, (p / ub) * 8, (p % ub) * 64);
  }
}

// Images of a data matrix (bae_upload_data): static scale 2^e
__global__ void __launch_bounds__(kThreads) bae_ximg_kernel(const float* __restrict__ src, int ld, int R, int ccK, int ccM, int e,
                                                           unsigned char* imgK, int kbsK, unsigned char* imgM, int kbsM) {
  const int ub = (((max(ccK, ccM) + 7) >> 3) + 63) >> 6;
  const int total = ((R + 7) >> 3) * ub;
  const float s = tc::exp2i(e);
  for (int item = blockIdx.x; item < total; item += gridDim.x)
    split_item_to_images(src, ld, R, ccK, ccM, s, imgK, kbsK, imgM, kbsM, (item / ub) * 8, (item % ub) * 64);
}

// Ones column (index d) of the MN-major image of a sigmoid layer's output, rows < `rows` of every group slot: 1.0 at the static
// scale 2^13 = FP16 0x7000 in the hi plane, 0 in the lo plane (the allocation's zeros). The epilogue that writes the image
// rewrites the unit that holds the column only when the layer's last 8 columns share it.
__global__ void bae_himg_ones_kernel(unsigned char* img, long long stride, int kbs, int d, int rows, int nslots) {
  const long long total = (long long)rows * nslots;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(i / rows), row = (int)(i % rows);
    unsigned char* p = img + (size_t)slot * stride + (size_t)(row >> 4) * kbs + (size_t)((row >> 3) & 1) * (size_t)(kbs >> 2) + (size_t)(d >> 3) * 128 +
                       (size_t)(row & 7) * 16 + (size_t)(d & 7) * 2;
    *reinterpret_cast<unsigned short*>(p) = 0x7000;
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d, training mode (MAIN:165): batch statistics forward, full backward
// ------------------------------------------------------------------------------------------------
// ---- BatchNorm phases for wide bottlenecks (d3 >= 16): a work item is (chain, 16 columns). 4 float4 column groups x
// 64 row lanes; every thread keeps 8 rows (8 or 16 float4 loads) in flight, because only 256 threads per SM run these
// phases and the tcgen05 kernels have no L1. The passes over a [rows x 16] slab (128 KB) after the first hit L2.
constexpr int kBnCols = 16, kBnLanes = kThreads / 4, kBnUnroll = 8;

// Sum of v over the 64 row lanes for each of this thread's 4 columns (result in every thread). red2: >= 128 doubles.
__device__ __forceinline__ void bn_colsum4(double (&v)[4], double* red2) {
  const int lane = TID & 31, w = TID >> 5, c4 = TID & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 4);
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
  }
  BLOCK_SYNC();
  if (lane < 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red2[w * kBnCols + c4 * 4 + k] = v[k];
  }
  BLOCK_SYNC();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += red2[i * kBnCols + c4 * 4 + k];
    v[k] = t;
  }
}

__device__ void run_bn_fwd_wide(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, double* red2) {
  const int d3 = S->d3, ld3 = S->ld3, N = ph.argRows, slot = ph.arg;
  const int nblk = (d3 + kBnCols - 1) / kBnCols;
  const int total = nblk * (c_end - c_begin);
  const int c4 = TID & 3, rl = TID >> 2;
  const int off_g = sh_seg[SEG_BN_W].pad_off, off_b = sh_seg[SEG_BN_BIAS].pad_off;
  const size_t slab = (size_t)S->n_max * ld3;
  const bool need = ph.needLang != 0;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / nblk;
    if (need && !S->ch.is_lang[chain]) continue;
    const int c0 = (item % nblk) * kBnCols + c4 * 4;
    const bool cv = c0 < d3;   // a float4 may reach past d3 (pad / ones column: read, never written)
    const float* __restrict__ Z = S->Z + (size_t)(chain - c_begin) * slab + c0;
    float* __restrict__ Y = S->Y + (size_t)(chain - c_begin) * slab + c0;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    if (cv)
      for (int base = rl; base < N; base += kBnLanes * kBnUnroll) {
        float4 z[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const int r = base + u * kBnLanes;
          z[u] = (r < N) ? *reinterpret_cast<const float4*>(Z + (size_t)r * ld3) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) { a[0] += (double)z[u].x; a[1] += (double)z[u].y; a[2] += (double)z[u].z; a[3] += (double)z[u].w; }
      }
    bn_colsum4(a, red2);
    const float mean[4] = {(float)(a[0] / N), (float)(a[1] / N), (float)(a[2] / N), (float)(a[3] / N)};
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    if (cv)
      for (int base = rl; base < N; base += kBnLanes * kBnUnroll) {
        float4 z[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const int r = base + u * kBnLanes;
          z[u] = (r < N) ? *reinterpret_cast<const float4*>(Z + (size_t)r * ld3)
                         : make_float4(mean[0], mean[1], mean[2], mean[3]);
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const float d0 = z[u].x - mean[0], d1 = z[u].y - mean[1], d2 = z[u].z - mean[2], d3_ = z[u].w - mean[3];
          q[0] += (double)d0 * (double)d0; q[1] += (double)d1 * (double)d1;
          q[2] += (double)d2 * (double)d2; q[3] += (double)d3_ * (double)d3_;
        }
      }
    bn_colsum4(q, red2);
    if (!cv) continue;
    const float* th = S->theta + (size_t)chain * S->Pp;
    float var[4], inv[4], gam[4], bet[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      var[k] = (float)(q[k] / N);
      inv[k] = 1.0f / sqrtf(var[k] + kBnEps);
      const bool ok = c0 + k < d3;
      gam[k] = ok ? th[off_g + c0 + k] : 0.f;
      bet[k] = ok ? th[off_b + c0 + k] : 0.f;
      if (rl == 0 && ok) {
        S->bn_mu[((size_t)slot * S->C_alloc + chain) * ld3 + c0 + k] = mean[k];
        S->bn_var[((size_t)slot * S->C_alloc + chain) * ld3 + c0 + k] = var[k];
        S->bn_inv[(size_t)chain * ld3 + c0 + k] = inv[k];
      }
    }
    const bool full4 = c0 + 4 <= d3;
    if (S->f16 && rl == 0) {
      // bound behind the FP16 scale of Y (3xFP16 split), without a pass over the data: |zhat| <= sqrt(N) for ANY batch
      // (var is the biased batch variance, so (z - mean)^2 <= N var), hence |y| <= |gamma| sqrt(N) + |beta|; at least 1,
      // the ones column next to Y (bias gradient of decode.1) being part of the operand
      const float rn = sqrtf((float)N);
      float bnd = 1.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) bnd = fmaxf(bnd, fmaf(fabsf(gam[k]), rn, fabsf(bet[k])));
      atomicMax(reinterpret_cast<unsigned*>(S->amax) + (size_t)AMAX_Y * S->C_alloc + chain, __float_as_uint(bnd));
    }
    for (int base = rl; base < N; base += kBnLanes * kBnUnroll) {
      float4 z[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int r = base + u * kBnLanes;
        if (r < N) z[u] = *reinterpret_cast<const float4*>(Z + (size_t)r * ld3);
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int r = base + u * kBnLanes;
        if (r >= N) continue;
        const float y[4] = {fmaf((z[u].x - mean[0]) * inv[0], gam[0], bet[0]), fmaf((z[u].y - mean[1]) * inv[1], gam[1], bet[1]),
                            fmaf((z[u].z - mean[2]) * inv[2], gam[2], bet[2]), fmaf((z[u].w - mean[3]) * inv[3], gam[3], bet[3])};
        float* yp = Y + (size_t)r * ld3;
        if (full4) {
          *reinterpret_cast<float4*>(yp) = make_float4(y[0], y[1], y[2], y[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < d3) yp[k] = y[k];
        }
      }
    }
  }
}

__device__ void run_bn_bwd_wide(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, double* red2) {
  const int d3 = S->d3, ld3 = S->ld3, N = ph.argRows, slot = ph.arg;
  const int nblk = (d3 + kBnCols - 1) / kBnCols;
  const int total = nblk * (c_end - c_begin);
  const int c4 = TID & 3, rl = TID >> 2;
  const int off_g = sh_seg[SEG_BN_W].pad_off, off_b = sh_seg[SEG_BN_BIAS].pad_off;
  const size_t slab = (size_t)S->n_max * ld3;
  const bool need = ph.needLang != 0;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / nblk;
    if (need && !S->ch.is_lang[chain]) continue;
    const int c0 = (item % nblk) * kBnCols + c4 * 4;
    const bool cv = c0 < d3;
    const float* __restrict__ Z = S->Z + (size_t)(chain - c_begin) * slab + c0;
    float* DY = S->DY + (size_t)(chain - c_begin) * slab + c0;
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, inv[4] = {0.f, 0.f, 0.f, 0.f}, gam[4] = {0.f, 0.f, 0.f, 0.f};
    if (cv) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c0 + k < d3) {
          mean[k] = S->bn_mu[((size_t)slot * S->C_alloc + chain) * ld3 + c0 + k];
          inv[k] = S->bn_inv[(size_t)chain * ld3 + c0 + k];
          gam[k] = S->theta[(size_t)chain * S->Pp + off_g + c0 + k];
        }
    }
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    if (cv)
      for (int base = rl; base < N; base += kBnLanes * kBnUnroll) {
        float4 dy[kBnUnroll], z[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const int r = base + u * kBnLanes;
          if (r < N) {
            dy[u] = *reinterpret_cast<const float4*>(DY + (size_t)r * ld3);
            z[u] = *reinterpret_cast<const float4*>(Z + (size_t)r * ld3);
          } else {
            dy[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            z[u] = dy[u];
          }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const float d[4] = {dy[u].x, dy[u].y, dy[u].z, dy[u].w};
          const float zz[4] = {z[u].x, z[u].y, z[u].z, z[u].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float zh = (zz[k] - mean[k]) * inv[k];
            s1[k] += (double)d[k];
            s2[k] += (double)d[k] * (double)zh;
          }
        }
      }
    bn_colsum4(s1, red2);
    bn_colsum4(s2, red2);
    if (!cv) continue;
    float* gr = S->grad + (size_t)chain * S->Pp;
    float m1[4], m2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (rl == 0 && c0 + k < d3) {
        gr[off_b + c0 + k] = (float)s1[k];   // d beta  = sum dy
        gr[off_g + c0 + k] = (float)s2[k];   // d gamma = sum dy * zhat
      }
      m1[k] = (float)(s1[k] / N) * gam[k];   // mean(dzhat), dzhat = dy * gamma
      m2[k] = (float)(s2[k] / N) * gam[k];   // mean(dzhat * zhat)
    }
    const bool full4 = c0 + 4 <= d3;
    if (S->f16 && rl == 0) {
      // bound behind the FP16 scale of the rewritten DY: |dz| <= inv (|gamma| max|dy| + |m1| + sqrt(N) |m2|), with max|dy|
      // the bound the GEMM epilogue that wrote dy folded into AMAX_DYIN
      const float din = S->amax[(size_t)AMAX_DYIN * S->C_alloc + chain], rn = sqrtf((float)N);
      float bnd = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) bnd = fmaxf(bnd, inv[k] * (fabsf(gam[k]) * din + fabsf(m1[k]) + rn * fabsf(m2[k])));
      if (bnd > 0.f) atomicMax(reinterpret_cast<unsigned*>(S->amax) + (size_t)AMAX_DY * S->C_alloc + chain, __float_as_uint(bnd));
    }
    for (int base = rl; base < N; base += kBnLanes * kBnUnroll) {
      float4 dy[kBnUnroll], z[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int r = base + u * kBnLanes;
        if (r < N) {
          dy[u] = *reinterpret_cast<const float4*>(DY + (size_t)r * ld3);
          z[u] = *reinterpret_cast<const float4*>(Z + (size_t)r * ld3);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int r = base + u * kBnLanes;
        if (r >= N) continue;
        const float d[4] = {dy[u].x, dy[u].y, dy[u].z, dy[u].w};
        const float zz[4] = {z[u].x, z[u].y, z[u].z, z[u].w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float zh = (zz[k] - mean[k]) * inv[k];
          o[k] = inv[k] * (d[k] * gam[k] - m1[k] - zh * m2[k]);
        }
        float* dp = DY + (size_t)r * ld3;
        if (full4) {
          *reinterpret_cast<float4*>(dp) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < d3) dp[k] = o[k];
        }
      }
    }
  }
}

__device__ void run_bn_fwd(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, double* red2 /*[8][32]*/) {
  const int d3 = S->d3, ld3 = S->ld3, N = ph.argRows, slot = ph.arg;
  // a CTA covers `cpb` columns with 256/cpb row lanes each (narrow bottlenecks such as Swiss Roll's 2 columns
  // still use every thread)
  int cpb = 32;
  while (cpb > 1 && (cpb >> 1) >= d3) cpb >>= 1;
  const int rln = kThreads / cpb;
  const int nblk = (d3 + cpb - 1) / cpb;
  const int total = nblk * (c_end - c_begin);
  const int cl = TID % cpb, rl = TID / cpb;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / nblk;
    if (ph.needLang && !S->ch.is_lang[chain]) continue;
    const int c = (item % nblk) * cpb + cl;
    const bool cv = c < d3;
    const float* Z = S->Z + (size_t)(chain - c_begin) * S->n_max * ld3;
    float* Y = S->Y + (size_t)(chain - c_begin) * S->n_max * ld3;
    double s = 0.0;
    if (cv)
      for (int r = rl; r < N; r += rln) s += (double)Z[(size_t)r * ld3 + c];
    BLOCK_SYNC();
    red2[rl * cpb + cl] = s;
    BLOCK_SYNC();
    double tot = 0.0;
    for (int i = 0; i < rln; ++i) tot += red2[i * cpb + cl];
    const float mean = (float)(tot / N);
    double q = 0.0;
    if (cv)
      for (int r = rl; r < N; r += rln) {
        float dz = Z[(size_t)r * ld3 + c] - mean;
        q += (double)dz * (double)dz;
      }
    BLOCK_SYNC();
    red2[rl * cpb + cl] = q;
    BLOCK_SYNC();
    double qt = 0.0;
    for (int i = 0; i < rln; ++i) qt += red2[i * cpb + cl];
    const float var = (float)(qt / N);
    const float inv = 1.0f / sqrtf(var + kBnEps);
    if (cv) {
      const float* th = S->theta + (size_t)chain * S->Pp;
      const float gam = th[S->seg[SEG_BN_W].pad_off + c];
      const float bet = th[S->seg[SEG_BN_BIAS].pad_off + c];
      if (rl == 0) {
        S->bn_mu[((size_t)slot * S->C_alloc + chain) * ld3 + c] = mean;
        S->bn_var[((size_t)slot * S->C_alloc + chain) * ld3 + c] = var;
        S->bn_inv[(size_t)chain * ld3 + c] = inv;
      }
      for (int r = rl; r < N; r += rln) {
        float zh = (Z[(size_t)r * ld3 + c] - mean) * inv;
        Y[(size_t)r * ld3 + c] = fmaf(zh, gam, bet);
      }
    }
  }
}

__device__ void run_bn_bwd(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, double* red2) {
  const int d3 = S->d3, ld3 = S->ld3, N = ph.argRows, slot = ph.arg;
  int cpb = 32;
  while (cpb > 1 && (cpb >> 1) >= d3) cpb >>= 1;
  const int rln = kThreads / cpb;
  const int nblk = (d3 + cpb - 1) / cpb;
  const int total = nblk * (c_end - c_begin);
  const int cl = TID % cpb, rl = TID / cpb;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / nblk;
    if (ph.needLang && !S->ch.is_lang[chain]) continue;
    const int c = (item % nblk) * cpb + cl;
    const bool cv = c < d3;
    const float* Z = S->Z + (size_t)(chain - c_begin) * S->n_max * ld3;
    float* DY = S->DY + (size_t)(chain - c_begin) * S->n_max * ld3;
    float mean = 0.f, inv = 0.f, gam = 0.f;
    if (cv) {
      mean = S->bn_mu[((size_t)slot * S->C_alloc + chain) * ld3 + c];
      inv = S->bn_inv[(size_t)chain * ld3 + c];
      gam = S->theta[(size_t)chain * S->Pp + S->seg[SEG_BN_W].pad_off + c];
    }
    double s1 = 0.0, s2 = 0.0;
    if (cv)
      for (int r = rl; r < N; r += rln) {
        float dy = DY[(size_t)r * ld3 + c];
        float zh = (Z[(size_t)r * ld3 + c] - mean) * inv;
        s1 += (double)dy;
        s2 += (double)dy * (double)zh;
      }
    BLOCK_SYNC();
    red2[rl * cpb + cl] = s1;
    red2[256 + rl * cpb + cl] = s2;
    BLOCK_SYNC();
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < rln; ++i) { t1 += red2[i * cpb + cl]; t2 += red2[256 + i * cpb + cl]; }
    if (cv) {
      float* gr = S->grad + (size_t)chain * S->Pp;
      if (rl == 0) {
        gr[S->seg[SEG_BN_BIAS].pad_off + c] = (float)t1;   // d beta  = sum dy
        gr[S->seg[SEG_BN_W].pad_off + c] = (float)t2;      // d gamma = sum dy * zhat
      }
      const float m1 = (float)(t1 / N) * gam;              // mean(dzhat), dzhat = dy * gamma
      const float m2 = (float)(t2 / N) * gam;              // mean(dzhat * zhat)
      for (int r = rl; r < N; r += rln) {
        float dy = DY[(size_t)r * ld3 + c];
        float zh = (Z[(size_t)r * ld3 + c] - mean) * inv;
        DY[(size_t)r * ld3 + c] = inv * (dy * gam - m1 - zh * m2);
      }
    }
  }
}

// Gradient of 4 consecutive entries: fixed-order sum of the K-split copies written by the weight-gradient GEMMs.
__device__ __forceinline__ float4 load_grad4(const float* __restrict__ gr, int p, int ks, size_t stride) {
  float4 g4 = *reinterpret_cast<const float4*>(&gr[p]);
  for (int s2 = 1; s2 < ks; ++s2) {
    const float4 t = *reinterpret_cast<const float4*>(&gr[p + s2 * stride]);
    g4.x += t.x; g4.y += t.y; g4.z += t.z; g4.w += t.w;
  }
  return g4;
}

// ------------------------------------------------------------------------------------------------
// Adam drift + proposal noise (first call, MAIN:473-474 / random walk MAIN:492) and the reverse
// drift (second call, MAIN:475-477), fused with the reductions the MH ratio needs.
// ------------------------------------------------------------------------------------------------
__device__ void run_adam(const DevState* S, int which, int c_begin, int c_end, double* red) {
  const int nit = S->n_adam_items;
  const int total = nit * (c_end - c_begin);
  const int Pp = S->Pp;
  const float step_w = S->step_w;
  const size_t gstride = (size_t)S->C_alloc * Pp;   // distance between the K-split copies of the gradient
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int chain = c_begin + item / nit;
    const int it = item % nit;
    const bool lang = S->ch.is_lang[chain] != 0;
    const bool alias = S->ch.alias[chain] != 0;
    double* part = S->adam_part + ((size_t)chain * nit + it) * 3;
    if (which == 2 && !lang) {
      if (TID == 0) part[2] = 0.0;
      continue;
    }
    const size_t base = (size_t)chain * Pp;
    float* __restrict__ th = S->theta + base;
    float* __restrict__ mm = S->m + base;
    float* __restrict__ vv = S->v + base;
    float* __restrict__ wc = S->wcur + base;
    const float* __restrict__ gr = S->grad + base;
    const int step = S->ch.step[chain];
    const int cg_id = S->ch.gid[chain];
    const float ss = (which == 1) ? S->ch.ss1[chain] : S->ch.ss2[chain];
    const float sq = (which == 1) ? S->ch.sq1[chain] : S->ch.sq2[chain];
    const bool plain = S->ch.plain[chain] != 0;               // SGD after pt / MALA: no Adam state involved
    const float da = S->ch.drift_a[chain], db = S->ch.drift_b[chain];
    // most 8192-float items lie inside one tensor: look the tensor up once per item then
    const int sg_first = find_seg(S, it * kAdamChunk), sg_last = find_seg(S, min(Pp, (it + 1) * kAdamChunk) - 4);
    const float inv_sq = 1.0f / sq;   // one IEEE division per item; per element: a multiply and an approximate reciprocal
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    float wmx = 0.f;   // 3xFP16 split: max |proposal| over the entries of weight MATRICES in this item (slot AMAX_W)
#pragma unroll 2
    for (int u = 0; u < kAdamChunk / (kThreads * 4); ++u) {
      const int p = it * kAdamChunk + (u * kThreads + TID) * 4;
      if (p >= Pp) continue;
      const int sg = (sg_first == sg_last) ? sg_first : find_seg(S, p);
      const SegLite sd = sh_seg[sg];
      if (!sd.trainable) continue;
      const int local = p - sd.pad_off;
      if (local >= sd.span) continue;   // alignment gap between tensors
      // local % ld without an integer division: local < 2^24, so the float quotient is off by at most one
      int col = local - __float2int_rz(__int2float_rz(local) * sd.inv_ld) * sd.ld;
      col += (col < 0) ? sd.ld : 0;
      col -= (col >= sd.ld) ? sd.ld : 0;
      const int nval = min(4, sd.cols - col);
      if (nval <= 0) continue;
      float4 t4 = *reinterpret_cast<const float4*>(&th[p]);
      float t[4] = {t4.x, t4.y, t4.z, t4.w};
      if (which == 1) {
        float o[4];
        if (lang && plain) {
          const float4 g4 = load_grad4(gr, p, sd.ksplit, gstride);
          const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = t[j] - da * g[j] - db * t[j];
          if (!alias) *reinterpret_cast<float4*>(&wc[p]) = t4;
        } else if (lang) {
          float4 g4 = load_grad4(gr, p, sd.ksplit, gstride);
          float4 m4 = *reinterpret_cast<const float4*>(&mm[p]);
          float4 v4 = *reinterpret_cast<const float4*>(&vv[p]);
          float g[4] = {g4.x, g4.y, g4.z, g4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
            v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
            float denom = fmaf(sqrtf(v[j]), inv_sq, kAdamEps);
            o[j] = t[j] - ss * __fdividef(m[j], denom);
          }
          *reinterpret_cast<float4*>(&mm[p]) = make_float4(m[0], m[1], m[2], m[3]);
          *reinterpret_cast<float4*>(&vv[p]) = make_float4(v[0], v[1], v[2], v[3]);
          if (!alias) *reinterpret_cast<float4*>(&wc[p]) = t4;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = t[j];
        }
        float nz[4];
        noise_quad(S, cg_id, step, sg, (unsigned)(local >> 2), nz);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nval) {
            float nj = nz[j] * step_w;
            o[j] = o[j] + nj;
            a0 = fmaf(nj, nj, a0);
            a1 = fmaf(o[j], o[j], a1);
          } else {
            o[j] = 0.f;
          }
        }
        if (sd.span > sd.ld) wmx = fmaxf(wmx, fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))));
        *reinterpret_cast<float4*>(&th[p]) = make_float4(o[0], o[1], o[2], o[3]);
      } else if (plain) {
        if (alias) continue;    // w IS the live model: w - w_prop_gd is exactly 0 (MAIN:476 with MAIN:449/555)
        const float4 g4 = load_grad4(gr, p, sd.ksplit, gstride);
        const float4 w4 = *reinterpret_cast<const float4*>(&wc[p]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, w0[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float wpg = t[j] - da * g[j] - db * t[j];   // w_prop_gd
          if (j < nval) {
            const float dl = w0[j] - wpg;
            a2 = fmaf(dl, dl, a2);
          }
        }
      } else {
        float4 g4 = load_grad4(gr, p, sd.ksplit, gstride);
        float4 m4 = *reinterpret_cast<const float4*>(&mm[p]);
        float4 v4 = *reinterpret_cast<const float4*>(&vv[p]);
        float g[4] = {g4.x, g4.y, g4.z, g4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
        float w0[4] = {0.f, 0.f, 0.f, 0.f};
        if (!alias) {
          float4 w4 = *reinterpret_cast<const float4*>(&wc[p]);
          w0[0] = w4.x; w0[1] = w4.y; w0[2] = w4.z; w0[3] = w4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
          v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
          float denom = fmaf(sqrtf(v[j]), inv_sq, kAdamEps);
          float wpg = t[j] - ss * __fdividef(m[j], denom);   // w_prop_gd
          if (!alias && j < nval) {
            float dl = w0[j] - wpg;
            a2 = fmaf(dl, dl, a2);
          }
        }
        *reinterpret_cast<float4*>(&mm[p]) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(&vv[p]) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    if (which == 1) {
      double s0 = block_sum_d((double)a0, red);
      double s1 = block_sum_d((double)a1, red);
      if (TID == 0) { part[0] = s0; part[1] = s1; }
      if (S->f16) fold_amax(S, AMAX_W, chain, wmx);   // zeroed by the launch in front of this phase
    } else {
      double s2 = block_sum_d((double)a2, red);
      if (TID == 0) part[2] = s2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Adam phases of the tcgen05 kernels: same arithmetic as run_adam, but the inputs (theta, the K-split copies of the
// gradient, m, v) are streamed into the idle GEMM stage buffers by the bulk-copy engine (cp.async.bulk + mbarrier,
// three 2048-float sub-chunks in flight per SM). Only 256 threads per SM run this phase and there is no L1; with the
// loads off the threads' registers the phase is bound by HBM instead of by load latency.
// ------------------------------------------------------------------------------------------------
constexpr int kAdamSub = 2048;                       // floats per staged sub-chunk (256 threads x 2 float4)
constexpr int kAdamArrays = 7;                       // theta, grad x 4 copies, m, v
constexpr int kAdamStages = 3;
constexpr int kAdamStageBytes = kAdamArrays * kAdamSub * 4;   // 56 KB
static_assert(kAdamStages * kAdamStageBytes <= tc::Smem::kTail, "Adam staging fits the GEMM tile ring + epilogue staging area");
static_assert(kAdamChunk % kAdamSub == 0, "sub-chunks tile a work item");

__device__ __forceinline__ void bulk_g2s(uint32_t dst_s, const void* src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
}

__device__ void run_adam_tc(const DevState* S, int which, int c_begin, int c_end, double* red, const tc::Smem& tsm) {
  const int nit = S->n_adam_items;
  const int total = nit * (c_end - c_begin);
  const int Pp = S->Pp;
  const float step_w = S->step_w;
  const size_t gstride = (size_t)S->C_alloc * Pp;   // distance between the K-split copies of the gradient
  const int ncopy = min(S->max_ksplit, 4);
  // every pointer and per-chain array base in a local: this phase has no L1, each S-> access is an L2 round trip
  float* const g_theta = S->theta; float* const g_m = S->m; float* const g_v = S->v; float* const g_wcur = S->wcur;
  const float* const g_grad = S->grad;
  const int* const a_lang = S->ch.is_lang; const int* const a_alias = S->ch.alias; const int* const a_step = S->ch.step;
  const float* const a_ss = (which == 1) ? S->ch.ss1 : S->ch.ss2;
  const float* const a_sq = (which == 1) ? S->ch.sq1 : S->ch.sq2;
  const int* const a_gid = S->ch.gid;
  double* const g_part = S->adam_part;
  constexpr int kSubs = kAdamChunk / kAdamSub;
  const uint32_t stage_s = tc::smem_u32(tsm.tiles);
  uint64_t* bars = tsm.spare_bars();                       // three spare mbarriers behind the GEMM ones
  const uint32_t bar_s = tc::smem_u32(bars);
  if (TID == 0) {
    for (int i = 0; i < kAdamStages; ++i) tc::mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  BLOCK_SYNC();
  // flat sequence of (item, sub-chunk) units of this CTA; unit u lives in stage u % 3
  const int my_items = (total > (int)blockIdx.x) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int units = my_items * kSubs;
  auto unit_item = [&](int u) { return (int)blockIdx.x + (u / kSubs) * (int)gridDim.x; };
  auto issue = [&](int u) {   // one thread: the bulk copies of unit u
    const int item = unit_item(u), sub = u % kSubs;
    const int chain = c_begin + item / nit, it = item % nit;
    const bool lang = a_lang[chain] != 0;
    const uint32_t st = stage_s + (u % kAdamStages) * kAdamStageBytes, bar = bar_s + (u % kAdamStages) * 8;
    const size_t off = (size_t)chain * Pp + (size_t)it * kAdamChunk + (size_t)sub * kAdamSub;
    const bool gmv = (which == 2) ? lang : lang;   // random-walk chains only need theta
    const int narr = gmv ? (3 + ncopy) : 1;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(narr * kAdamSub * 4)) : "memory");
    bulk_g2s(st, g_theta + off, kAdamSub * 4, bar);
    if (gmv) {
      bulk_g2s(st + 1 * kAdamSub * 4, g_m + off, kAdamSub * 4, bar);
      bulk_g2s(st + 2 * kAdamSub * 4, g_v + off, kAdamSub * 4, bar);
      for (int k = 0; k < ncopy; ++k) bulk_g2s(st + (3 + k) * kAdamSub * 4, g_grad + off + k * gstride, kAdamSub * 4, bar);
    }
  };
  if (TID == 0)
    for (int u = 0; u < min(units, kAdamStages); ++u) issue(u);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  bool lang = false, alias = false;
  int step = 0;
  float ss = 0.f, sq = 1.f;
  for (int u = 0; u < units; ++u) {
    const int item = unit_item(u), sub = u % kSubs;
    const int chain = c_begin + item / nit, it = item % nit;
    if (sub == 0) {   // per-item scalars: one round of loads per 4 staged sub-chunks
      lang = a_lang[chain] != 0; alias = a_alias[chain] != 0; step = a_step[chain]; ss = a_ss[chain]; sq = a_sq[chain];
    }
    const int stg = u % kAdamStages;
    tc::mbar_wait(&bars[stg], (uint32_t)((u / kAdamStages) & 1));
    const bool active = !(which == 2 && !lang);
    if (active) {
      const size_t base = (size_t)chain * Pp;
      float* __restrict__ th = g_theta + base;
      float* __restrict__ mm = g_m + base;
      float* __restrict__ vv = g_v + base;
      float* __restrict__ wc = g_wcur + base;
      const int cg_id = a_gid[chain];
      const uint32_t sb = stage_s + stg * kAdamStageBytes;
#pragma unroll
      for (int e = 0; e < kAdamSub / (kThreads * 4); ++e) {
        const int q = (e * kThreads + TID) * 4;                      // float index inside the sub-chunk
        const int p = it * kAdamChunk + sub * kAdamSub + q;
        const int sg = find_seg(S, p);
        const SegLite sd = sh_seg[sg];
        const int local = p - sd.pad_off;
        if (!sd.trainable || local >= sd.span) continue;             // BatchNorm buffer, or the alignment gap between tensors
        const int nval = (sd.cols == sd.ld) ? 4 : min(4, sd.cols - local % sd.ld);
        if (nval <= 0) continue;
        const float4 t4 = tc::lds128(sb + q * 4);
        float t[4] = {t4.x, t4.y, t4.z, t4.w};
        float g[4] = {0.f, 0.f, 0.f, 0.f}, m[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
        if (lang) {
          const float4 m4 = tc::lds128(sb + (1 * kAdamSub + q) * 4), v4 = tc::lds128(sb + (2 * kAdamSub + q) * 4);
          m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
          v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
          const int ks = sd.ksplit;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < ks) {
              const float4 x = tc::lds128(sb + ((3 + k) * kAdamSub + q) * 4);
              g[0] += x.x; g[1] += x.y; g[2] += x.z; g[3] += x.w;
            }
          for (int k = 4; k < ks; ++k) {   // more than four K splits (very long data sets): the rest straight from HBM
            const float4 x = *reinterpret_cast<const float4*>(g_grad + base + p + k * gstride);
            g[0] += x.x; g[1] += x.y; g[2] += x.z; g[3] += x.w;
          }
        }
        if (which == 1) {
          float o[4];
          if (lang) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
              v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
              float denom = fmaf(sqrtf(v[j]), 1.0f / sq, kAdamEps);   // same arithmetic as run_adam (the compiler hoists the division)
              o[j] = t[j] - ss * __fdividef(m[j], denom);
            }
            *reinterpret_cast<float4*>(&mm[p]) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(&vv[p]) = make_float4(v[0], v[1], v[2], v[3]);
            if (!alias) *reinterpret_cast<float4*>(&wc[p]) = t4;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = t[j];
          }
          float nz[4];
          noise_quad(S, cg_id, step, sg, (unsigned)(local >> 2), nz);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nval) {
              float nj = nz[j] * step_w;
              o[j] = o[j] + nj;
              a0 = fmaf(nj, nj, a0);
              a1 = fmaf(o[j], o[j], a1);
            } else {
              o[j] = 0.f;
            }
          }
          *reinterpret_cast<float4*>(&th[p]) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          float w0[4] = {0.f, 0.f, 0.f, 0.f};
          if (!alias) {
            const float4 w4 = *reinterpret_cast<const float4*>(&wc[p]);
            w0[0] = w4.x; w0[1] = w4.y; w0[2] = w4.z; w0[3] = w4.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
            v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
            float denom = fmaf(sqrtf(v[j]), 1.0f / sq, kAdamEps);   // same arithmetic as run_adam (the compiler hoists the division)
            float wpg = t[j] - ss * __fdividef(m[j], denom);   // w_prop_gd
            if (!alias && j < nval) {
              float dl = w0[j] - wpg;
              a2 = fmaf(dl, dl, a2);
            }
          }
          *reinterpret_cast<float4*>(&mm[p]) = make_float4(m[0], m[1], m[2], m[3]);
          *reinterpret_cast<float4*>(&vv[p]) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    // every thread is done reading the stage: refill it; at the end of an item, reduce its partial sums
    BLOCK_SYNC();
    if (TID == 0 && u + kAdamStages < units) issue(u + kAdamStages);
    if (sub == kSubs - 1) {
      double* part = g_part + ((size_t)chain * nit + it) * 3;
      if (which == 1) {
        double s0 = block_sum_d((double)a0, red);
        double s1 = block_sum_d((double)a1, red);
        if (TID == 0) { part[0] = s0; part[1] = s1; }
      } else {
        double s2 = active ? block_sum_d((double)a2, red) : 0.0;
        if (TID == 0) part[2] = s2;
      }
      a0 = a1 = a2 = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scalar prologue: branch select, tempered -> canonical switch, draws        MAIN:435-452, 500, 526
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double loglik_d(double sse, double n, double tau_sq, double temp) {
  // MAIN:325-329: tau_sq MULTIPLIES the SSE, as written in the reference.
  return (-(n * 0.5) * log(2.0 * CUDART_PI * tau_sq) - 0.5 * tau_sq * sse) / temp;
}

__device__ void step_scalars(const DevState* S, int chain_global, int step, double& lx, double& eta_noise, double& u) {
  u4 c;
  c.x = 0u;
  c.y = 0xFFu | (STREAM_STEP_SCALARS << 8);
  c.z = (unsigned)step;
  c.w = (unsigned)chain_global;
  u4 r = philox4x32_10(c, S->seed_lo, S->seed_hi);
  lx = (double)u01(r.x);
  float n0, n1;
  box_muller(u01(r.y), u01(r.z), n0, n1);
  eta_noise = (double)n0 * (double)S->step_eta;
  u = (double)u01(r.w);
}

__device__ void run_prologue(const DevState* S, int c_begin, int c_end) {
  const int gtid = blockIdx.x * kThreads + TID;
  for (int chain = c_begin + gtid; chain < c_end; chain += gridDim.x * kThreads) {
    const ChainArrays& ch = S->ch;
    const int i = ch.step[chain];
    if (i < S->pt_step) ch.adapt[chain] = ch.T[chain];                        // MAIN:437-438
    if (i == S->pt_step && ch.init_count[chain] == 0) {                         // MAIN:440-445
      ch.adapt[chain] = 1.0;
      // evaluated on w_proposal, the chain's OWN last proposal -- it does not travel with exchanges
      ch.lik[chain] = loglik_d(ch.prop_sse[chain], (double)S->n_train * S->d0, 1.0, 1.0);
      ch.init_count[chain] = 1;
    }
    double lx, en, u;
    step_scalars(S, ch.gid[chain], i, lx, en, u);
    ch.lx[chain] = lx;
    ch.eta_noise[chain] = en;
    ch.u[chain] = u;
    const int lang = (S->use_langevin && lx < (double)S->l_prob) ? 1 : 0;       // MAIN:452
    ch.is_lang[chain] = lang;
    // drift of this iteration: Adam (the reference's setting), or a plain gradient step w - a * grad - b * w:
    //   drift_mode 1 (experiment_type = 1, MAIN:454-470): i > pt -> fresh SGD(lr = 0.1): a = 0.1, b = 0
    //   drift_mode 2 (MALA): a = (eps^2 / 2) * tau * n / (2 T), b = (eps^2 / 2) / sigma^2 with grad = d MSE / d w
    int plain = 0;
    float da = 0.f, db = 0.f;
    if (S->drift_mode == 1 && i > S->pt_step) { plain = 1; da = 0.1f; }
    if (S->drift_mode == 2) {
      plain = 1;
      const double h2 = 0.5 * (double)S->step_w * (double)S->step_w;
      da = (float)(h2 * 0.5 * exp(ch.eta[chain]) * ((double)S->n_train * S->d0) / ch.adapt[chain]);
      db = (float)(h2 / (double)S->sigma_sq);
    }
    ch.plain[chain] = plain;
    ch.drift_a[chain] = da;
    ch.drift_b[chain] = db;
    if (lang && !plain) {
      const int t1 = ch.adam_t[chain] + 1, t2 = t1 + 1;
      ch.ss1[chain] = (float)((double)S->lr / (1.0 - pow((double)kB1, (double)t1)));
      ch.sq1[chain] = (float)sqrt(1.0 - pow((double)kB2, (double)t1));
      ch.ss2[chain] = (float)((double)S->lr / (1.0 - pow((double)kB1, (double)t2)));
      ch.sq2[chain] = (float)sqrt(1.0 - pow((double)kB2, (double)t2));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scalar epilogue: BatchNorm-buffer bookkeeping, likelihood, prior, reverse-proposal density, MH test
// (one CTA per chain)                                                              MAIN:476-592
// ------------------------------------------------------------------------------------------------
__device__ void run_epilogue(const DevState* S, int c_begin, int c_end, double* red) {
  const int d3 = S->d3, ld3 = S->ld3;
  const ChainArrays& ch = S->ch;
  for (int chain = c_begin + blockIdx.x; chain < c_end; chain += gridDim.x) {
    const bool lang = ch.is_lang[chain] != 0;
    const bool alias = ch.alias[chain] != 0;
    const int step = ch.step[chain];
    const int cg_id = ch.gid[chain];
    float* th = S->theta + (size_t)chain * S->Pp;
    float* wb = S->wb + (size_t)chain * S->nbuf;   // rm[d3], rv[d3], nbt
    float* rm = th + S->seg[SEG_BN_RM].pad_off;
    float* rv = th + S->seg[SEG_BN_RV].pad_off;
    float* nbt_live = th + S->seg[SEG_BN_NBT].pad_off;
    const float* mu1 = S->bn_mu + ((size_t)0 * S->C_alloc + chain) * ld3;
    const float* va1 = S->bn_var + ((size_t)0 * S->C_alloc + chain) * ld3;
    const float* mu2 = S->bn_mu + ((size_t)1 * S->C_alloc + chain) * ld3;
    const float* va2 = S->bn_var + ((size_t)1 * S->C_alloc + chain) * ld3;
    const float* mut = S->bn_mu + ((size_t)2 * S->C_alloc + chain) * ld3;
    const float* vat = S->bn_var + ((size_t)2 * S->C_alloc + chain) * ld3;
    const float ub_tr = (float)S->n_train / (float)(S->n_train - 1);
    const float ub_te = (float)S->n_test / (float)(S->n_test - 1);

    // --- gather partial sums (fixed order -> deterministic)
    double p_sse_tr = 0.0, p_sse_te = 0.0, p_noise = 0.0, p_prop = 0.0, p_wc = 0.0;
    const int nl_tr = S->loss_stride;
    const double* lp_tr = S->loss_part + (((size_t)1 * S->C_alloc + chain) * S->loss_stride) * 2;
    const double* lp_te = S->loss_part + (((size_t)2 * S->C_alloc + chain) * S->loss_stride) * 2;
    for (int i = TID; i < nl_tr; i += kThreads) { p_sse_tr += lp_tr[i * 2]; p_sse_te += lp_te[i * 2]; }
    const double* ap = S->adam_part + (size_t)chain * S->n_adam_items * 3;
    for (int i = TID; i < S->n_adam_items; i += kThreads) {
      p_noise += ap[i * 3 + 0];
      p_prop += ap[i * 3 + 1];
      if (lang) p_wc += ap[i * 3 + 2];
    }
    // --- BatchNorm buffer bookkeeping (all sorted-key buffer entries take noise, MAIN:259-260)
    double b_wc = 0.0, b_wp = 0.0, b_prior = 0.0;
    for (int c = TID; c < d3; c += kThreads) {
      float nzm[4], nzv[4];
      noise_quad(S, cg_id, step, SEG_BN_RM, (unsigned)(c >> 2), nzm);
      noise_quad(S, cg_id, step, SEG_BN_RV, (unsigned)(c >> 2), nzv);
      const float n_rm = nzm[c & 3] * S->step_w, n_rv = nzv[c & 3] * S->step_w;
      const float base_rm = alias ? rm[c] : wb[c];
      const float base_rv = alias ? rv[c] : wb[d3 + c];
      float p_rm, p_rv;
      if (lang) {
        // first drift forward at w (MAIN:473): running stats of the loaded dict are updated
        const float g_rm = (1.0f - kBnMom) * base_rm + kBnMom * mu1[c];
        const float g_rv = (1.0f - kBnMom) * base_rv + kBnMom * va1[c] * ub_tr;
        p_rm = g_rm + n_rm;                                  // w_proposal buffers (MAIN:474)
        p_rv = g_rv + n_rv;
        if (!alias) {
          // second drift forward at w_proposal (MAIN:475) -> w_prop_gd buffers
          const float q_rm = (1.0f - kBnMom) * p_rm + kBnMom * mu2[c];
          const float q_rv = (1.0f - kBnMom) * p_rv + kBnMom * va2[c] * ub_tr;
          const double d1 = (double)wb[c] - (double)q_rm, d2 = (double)wb[d3 + c] - (double)q_rv;
          b_wc += d1 * d1 + d2 * d2;
        }
      } else {
        p_rm = base_rm + n_rm;                               // MAIN:492
        p_rv = base_rv + n_rv;
      }
      b_wp += (double)n_rm * n_rm + (double)n_rv * n_rv;
      b_prior += (double)p_rm * p_rm + (double)p_rv * p_rv;
      // likelihood forwards reload w_proposal (MAIN:505, 508); the test forward is the last to touch the live model
      rm[c] = (1.0f - kBnMom) * p_rm + kBnMom * mut[c];
      rv[c] = (1.0f - kBnMom) * p_rv + kBnMom * vat[c] * ub_te;
      // stash proposal buffers; they become `w`'s buffers on accept
      S->wb_tmp[(size_t)chain * S->nbuf + c] = p_rm;
      S->wb_tmp[(size_t)chain * S->nbuf + d3 + c] = p_rv;
    }
    const double sse_tr = block_sum_d(p_sse_tr, red);
    const double sse_te = block_sum_d(p_sse_te, red);
    const double s_noise = block_sum_d(p_noise, red);
    const double s_prop = block_sum_d(p_prop, red);
    const double s_wc = block_sum_d(p_wc, red);
    const double sb_wc = block_sum_d(b_wc, red);
    const double sb_wp = block_sum_d(b_wp, red);
    const double sb_prior = block_sum_d(b_prior, red);
    BLOCK_SYNC();
    if (TID == 0) {
      // --- num_batches_tracked (int64 in the model, float in the dicts)
      float nzn[4];
      noise_quad(S, cg_id, step, SEG_BN_NBT, 0u, nzn);
      const float n_nbt = nzn[0] * S->step_w;
      const float base_nbt = alias ? nbt_live[0] : wb[2 * d3];
      float p_nbt;
      double nbt_wc = 0.0, nbt_wp = 0.0;
      if (lang) {
        const float g_nbt = truncf(base_nbt) + 1.0f;          // load (trunc) + first forward
        p_nbt = g_nbt + n_nbt;                                // int64 + float32 -> float32 (MAIN:260)
        nbt_wp = (double)p_nbt - (double)g_nbt;
        if (!alias) {
          const float q_nbt = truncf(p_nbt) + 1.0f;           // load + second forward
          nbt_wc = (double)wb[2 * d3] - (double)q_nbt;
        }
      } else {
        p_nbt = base_nbt + n_nbt;
        nbt_wp = (double)p_nbt - (double)base_nbt;
      }
      nbt_live[0] = truncf(p_nbt) + 1.0f;
      S->wb_tmp[(size_t)chain * S->nbuf + 2 * d3] = p_nbt;

      const double sigma_sq = (double)S->step_w * (double)S->step_w;
      double diff_prop = 0.0;
      if (lang) {
        const double first = -0.5 * (s_wc + sb_wc + nbt_wc * nbt_wc) / sigma_sq;        // MAIN:481
        const double second = -0.5 * (s_noise + sb_wp + nbt_wp * nbt_wp) / sigma_sq;    // MAIN:482
        diff_prop = first - second;
      }
      const double eta_pro = ch.eta[chain] + ch.eta_noise[chain];                         // MAIN:500
      const double tau_pro = exp(eta_pro);
      const double adapt = ch.adapt[chain];
      const double n_tr = (double)S->n_train * S->d0, n_te = (double)S->n_test * S->d0;
      const double lik_prop = loglik_d(sse_tr, n_tr, tau_pro, adapt);                     // MAIN:505
      const double mse_tr = sse_tr / n_tr, mse_te = sse_te / n_te;
      const double sumsq = s_prop + sb_prior + (double)p_nbt * (double)p_nbt;
      const double prior_prop = -((double)S->w_size * 0.5) * log((double)S->sigma_sq)     // MAIN:331-336
                                - sumsq / (2.0 * (double)S->sigma_sq)
                                - (1.0 + (double)S->nu1) * log(tau_pro) - ((double)S->nu2 / tau_pro);
      const double lik_cur = ch.lik[chain];
      const double sum_value = (lik_prop - lik_cur) + (prior_prop - ch.prior_cur[chain]) + diff_prop;  // MAIN:524
      const double log_u = log(ch.u[chain]);                                              // MAIN:526
      const bool accept = log_u < sum_value;                                              // MAIN:532
      if (accept) {
        ch.nacc[chain] += 1;
        ch.lik[chain] = lik_prop;
        ch.prior_cur[chain] = prior_prop;
        ch.eta[chain] = eta_pro;
        ch.mse_tr[chain] = mse_tr;
        ch.mse_te[chain] = mse_te;
        ch.alias[chain] = 0;                                  // w = deepcopy(w_proposal)   MAIN:537
      } else {
        ch.alias[chain] = 1;                                  // w = old_w (live tensors)   MAIN:555
      }
      if (lang) {
        ch.nlang[chain] += 1;
        if (!ch.plain[chain]) ch.adam_t[chain] += 2;
      }
      ch.last_sse[chain] = sse_tr;
      ch.prop_sse[chain] = sse_tr;
      if (chain < S->C && step < S->samples) {
        bae_trace& tr = S->traces[(size_t)chain * S->samples + step];
        tr.likelihood_proposal = lik_prop;
        tr.likelihood = lik_cur;
        tr.sum_value = sum_value;
        tr.log_u = log_u;
        tr.prior_proposal = prior_prop;
        tr.diff_prop = diff_prop;
        tr.mse_train = ch.mse_tr[chain];
        tr.mse_test = ch.mse_te[chain];
        tr.mse_train_proposal = mse_tr;
        tr.mse_test_proposal = mse_te;
        tr.eta = ch.eta[chain];
        tr.accepted = accept ? 1 : 0;
        tr.langevin = lang ? 1 : 0;
        tr.step = step;
        for (int k = 0; k < 13; ++k) tr.weights[k] = S->trace_idx_pad[k] >= 0 ? th[S->trace_idx_pad[k]] : 0.f;
      }
      ch.step[chain] = step + 1;
      red[10] = accept ? 1.0 : 0.0;
    }
    BLOCK_SYNC();
    if (red[10] != 0.0) {   // accept: `w` buffers <- proposal buffers
      for (int c = TID; c < S->nbuf; c += kThreads) wb[c] = S->wb_tmp[(size_t)chain * S->nbuf + c];
    }
    BLOCK_SYNC();
  }
}

// ------------------------------------------------------------------------------------------------
// Replica exchange: sequential adjacent-pair sweep per ensemble            MAIN:962-1011, 1059-1065
// ------------------------------------------------------------------------------------------------
__device__ double swap_uniform(const DevState* S, int ensemble_global, int round, int pair) {
  u4 c;
  c.x = (unsigned)pair;
  c.y = 0xFEu | (STREAM_SWAP << 8);
  c.z = (unsigned)round;
  c.w = (unsigned)ensemble_global;
  u4 r = philox4x32_10(c, S->seed_lo, S->seed_hi);
  return (double)u01(r.x);
}

__device__ void run_swap_decide(const DevState* S) {
  const int n = S->n_rungs;
  const int nens = S->C / n;
  const int gtid = blockIdx.x * kThreads + TID;
  const ChainArrays& ch = S->ch;
  const double ntr = (double)S->n_train * S->d0;
  for (int e = gtid; e < nens; e += gridDim.x * kThreads) {
    const int b = e * n;
    // queue slots: configuration id, and the (lhood, T) slots that travel with the message (MAIN:999-1005)
    for (int p = 0; p < n; ++p) S->swap_src[b + p] = b + p;
    double lh_prev = ch.lik[b], T_prev = ch.adapt[b];
    int cfg_prev = b;
    for (int p = 0; p + 1 < n; ++p) {
      const int cfg1 = cfg_prev, cfg2 = b + p + 1;
      const double lhood1 = lh_prev, T1 = T_prev;
      const double lhood2 = ch.lik[cfg2], T2 = ch.adapt[cfg2];
      const double lhood12 = loglik_d(ch.last_sse[cfg1], ntr, exp(ch.eta[cfg1]), T2);   // MAIN:980
      const double lhood21 = loglik_d(ch.last_sse[cfg2], ntr, exp(ch.eta[cfg2]), T1);   // MAIN:982
      const double ex = (lhood12 - lhood1) + (lhood21 - lhood2);
      const double prob = fmin(1.0, exp(ex));                                            // MAIN:988
      const int round = ch.step[b] / S->swap_interval - 1;
      const double u = swap_uniform(S, (S->chain_id_base / n) + e, round, p);
      const bool sw = u < prob;
      atomicAdd((unsigned long long*)S->total_swap, 1ull);
      if (sw) {
        atomicAdd((unsigned long long*)S->num_swap, 1ull);
        S->swap_src[b + p] = cfg2;       // position p receives configuration 2
        cfg_prev = cfg1;                 // configuration 1 moves on to position p+1
        lh_prev = lhood12;               // and carries lhood12 / T1 in its message slots
        T_prev = T1;
      } else {
        S->swap_src[b + p] = cfg1;
        cfg_prev = cfg2;
        lh_prev = lhood2;
        T_prev = T2;
      }
    }
    S->swap_src[b + n - 1] = cfg_prev;
  }
}

// gather: wcur[p] <- theta[src[p]] (trainable + buffers view of `w`), scalars to tmp
__device__ void run_swap_gather(const DevState* S) {
  const int nit = (S->Pp + kAdamChunk - 1) / kAdamChunk;
  const int total = nit * S->C;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int p = item / nit, it = item % nit;
    const int src = S->swap_src[p];
    if (it == 0) {
      // buffers of the posted `w`: detached copy or the live model's buffers (MAIN:595)
      const float* sth = S->theta + (size_t)src * S->Pp;
      const bool al = S->ch.alias[src] != 0;
      for (int c = TID; c < S->nbuf; c += kThreads) {
        float v;
        if (al) {
          if (c < S->d3) v = sth[S->seg[SEG_BN_RM].pad_off + c];
          else if (c < 2 * S->d3) v = sth[S->seg[SEG_BN_RV].pad_off + c - S->d3];
          else v = sth[S->seg[SEG_BN_NBT].pad_off];
        } else {
          v = S->wb[(size_t)src * S->nbuf + c];
        }
        S->wb_tmp[(size_t)p * S->nbuf + c] = v;
      }
      if (TID == 0) {
        S->swap_tmp[p * 2 + 0] = S->ch.eta[src];
        S->swap_tmp[p * 2 + 1] = S->ch.last_sse[src];
      }
    }
    if (src == p) continue;
    const float4* s4 = reinterpret_cast<const float4*>(S->theta + (size_t)src * S->Pp + (size_t)it * kAdamChunk);
    float4* d4 = reinterpret_cast<float4*>(S->wcur + (size_t)p * S->Pp + (size_t)it * kAdamChunk);
    const int nv = min(kAdamChunk, S->Pp - it * kAdamChunk) / 4;
    for (int i = TID; i < nv; i += kThreads) d4[i] = s4[i];
  }
}

__device__ void run_swap_scatter(const DevState* S) {
  const int nit = (S->Pp + kAdamChunk - 1) / kAdamChunk;
  const int total = nit * S->C;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int p = item / nit, it = item % nit;
    const int src = S->swap_src[p];
    if (it == 0) {
      for (int c = TID; c < S->nbuf; c += kThreads) S->wb[(size_t)p * S->nbuf + c] = S->wb_tmp[(size_t)p * S->nbuf + c];
      if (TID == 0) {
        S->ch.eta[p] = S->swap_tmp[p * 2 + 0];       // MAIN:603
        S->ch.last_sse[p] = S->swap_tmp[p * 2 + 1];
        S->ch.alias[p] = 0;                          // MAIN:602: w = dictfromlist(...) is a detached dict
      }
    }
    if (src == p) continue;
    // only trainable entries move into the live model; its BatchNorm buffers are reloaded from `w` next step
    const float4* s4 = reinterpret_cast<const float4*>(S->wcur + (size_t)p * S->Pp + (size_t)it * kAdamChunk);
    float* dbase = S->theta + (size_t)p * S->Pp;
    const int nv = min(kAdamChunk, S->Pp - it * kAdamChunk) / 4;
    for (int i = TID; i < nv; i += kThreads) {
      const int pidx = it * kAdamChunk + i * 4;
      const int sg = find_seg(S, pidx);
      if (!S->seg[sg].trainable) continue;
      *reinterpret_cast<float4*>(dbase + pidx) = s4[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Exchange round of a ladder SHARDED over ranks (one process per GPU; bae_swap_exchange): every rank holds rungs
// [rank L, rank L + L) of each of its E ensembles (local chain = e L + l), R = L x world rungs per ladder.
//   xtable  : (eta, stored likelihood, adapttemp, SSE of w, alias) of the resident chains -> xt_local, all-gathered by NCCL
//   xsweep  : the same sequential sweep as run_swap_decide on the all-gathered table; every rank computes the identical
//             permutation (same doubles, same Philox uniforms), so no second collective is needed for agreement
//   xgather : same-GPU moves into the wcur staging slabs + the posted BatchNorm buffers of every position
//   (NCCL)  : configurations whose destination is on another rank: theta slab -> the receiver's wcur slab, wb -> wb_tmp
//   xscatter: staged configurations into the live models, eta / SSE from the table, `w` becomes a detached dict (MAIN:602)
// ------------------------------------------------------------------------------------------------
constexpr int kXtCols = 5;

__device__ __forceinline__ double xt_at(const DevState* S, int L, int e, int rung, int k) {
  return S->xt_all[((size_t)(rung / L) * S->C + (size_t)e * L + (rung % L)) * kXtCols + k];
}

__global__ void bae_xtable_kernel(const DevState* __restrict__ S) {
  const ChainArrays& ch = S->ch;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < S->C; c += gridDim.x * blockDim.x) {
    double* t = S->xt_local + (size_t)c * kXtCols;
    t[0] = ch.eta[c]; t[1] = ch.lik[c]; t[2] = ch.adapt[c]; t[3] = ch.last_sse[c]; t[4] = (double)ch.alias[c];
  }
}

__global__ void bae_xsweep_kernel(const DevState* __restrict__ S) {
  const int L = S->n_rungs, R = L * S->shard_world, E = S->C / L;
  const double ntr = (double)S->n_train * S->d0;
  const int ens_base = S->chain_id_base / R;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    int* src = S->xsrc + (size_t)e * R;
    double lh_prev = xt_at(S, L, e, 0, 1), T_prev = xt_at(S, L, e, 0, 2);
    int cfg_prev = 0;
    const int round = S->ch.step[e * L] / S->swap_interval - 1;
    long long nsw = 0;
    for (int p = 0; p + 1 < R; ++p) {
      const int cfg1 = cfg_prev, cfg2 = p + 1;
      const double lhood1 = lh_prev, T1 = T_prev;
      const double lhood2 = xt_at(S, L, e, cfg2, 1), T2 = xt_at(S, L, e, cfg2, 2);
      const double lhood12 = loglik_d(xt_at(S, L, e, cfg1, 3), ntr, exp(xt_at(S, L, e, cfg1, 0)), T2);   // MAIN:980
      const double lhood21 = loglik_d(xt_at(S, L, e, cfg2, 3), ntr, exp(xt_at(S, L, e, cfg2, 0)), T1);   // MAIN:982
      const double ex = (lhood12 - lhood1) + (lhood21 - lhood2);
      const double prob = fmin(1.0, exp(ex));                                                            // MAIN:988
      const double u = swap_uniform(S, ens_base + e, round, p);
      if (u < prob) {
        ++nsw;
        src[p] = cfg2; cfg_prev = cfg1; lh_prev = lhood12; T_prev = T1;
      } else {
        src[p] = cfg1; cfg_prev = cfg2; lh_prev = lhood2; T_prev = T2;
      }
    }
    src[R - 1] = cfg_prev;
    atomicAdd((unsigned long long*)S->total_swap, (unsigned long long)(R - 1));
    atomicAdd((unsigned long long*)S->num_swap, (unsigned long long)nsw);
  }
}

// posted BatchNorm buffers of local chain `c` (rm[d3], rv[d3], nbt): the detached copy, or the live model's (MAIN:595)
__device__ __forceinline__ float posted_buffer(const DevState* S, int c, int i) {
  if (S->ch.alias[c] == 0) return S->wb[(size_t)c * S->nbuf + i];
  const float* th = S->theta + (size_t)c * S->Pp;
  if (i < S->d3) return th[S->seg[SEG_BN_RM].pad_off + i];
  if (i < 2 * S->d3) return th[S->seg[SEG_BN_RV].pad_off + i - S->d3];
  return th[S->seg[SEG_BN_NBT].pad_off];
}

__global__ void __launch_bounds__(kThreads) bae_xgather_kernel(const DevState* __restrict__ S) {
  const int L = S->n_rungs, R = L * S->shard_world, rank = S->shard_rank;
  const int nit = S->Pp / kAdamChunk;
  const int total = nit * S->C;
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int p = item / nit, it = item % nit;
    const int e = p / L, rung = rank * L + p % L;
    const int sr = S->xsrc[(size_t)e * R + rung];
    if (sr / L != rank) continue;                 // arrives over NCCL
    const int src = e * L + sr % L;
    if (it == 0)
      for (int i = threadIdx.x; i < S->nbuf; i += kThreads) S->wb_tmp[(size_t)p * S->nbuf + i] = posted_buffer(S, src, i);
    if (src == p) continue;
    const float4* s4 = reinterpret_cast<const float4*>(S->theta + (size_t)src * S->Pp + (size_t)it * kAdamChunk);
    float4* d4 = reinterpret_cast<float4*>(S->wcur + (size_t)p * S->Pp + (size_t)it * kAdamChunk);
    for (int i = threadIdx.x; i < kAdamChunk / 4; i += kThreads) d4[i] = s4[i];
  }
}

__global__ void __launch_bounds__(kThreads) bae_xscatter_kernel(const DevState* __restrict__ S) {
  const int L = S->n_rungs, R = L * S->shard_world, rank = S->shard_rank;
  const int nit = S->Pp / kAdamChunk;
  const int total = nit * S->C;
  __shared__ int seg_off[kNumSegs], seg_train[kNumSegs];
  for (int i = threadIdx.x; i < kNumSegs; i += kThreads) { seg_off[i] = S->seg[i].pad_off; seg_train[i] = S->seg[i].trainable; }
  __syncthreads();
  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int p = item / nit, it = item % nit;
    const int e = p / L, rung = rank * L + p % L;
    const int sr = S->xsrc[(size_t)e * R + rung];
    const bool remote = sr / L != rank;
    const float* stage = S->wcur + (size_t)p * S->Pp;
    if (it == 0) {
      // buffers of the received dict: a sender whose `w` was the live model posted them inside the theta slab
      const bool from_slab = remote && xt_at(S, L, e, sr, 4) != 0.0;
      for (int i = threadIdx.x; i < S->nbuf; i += kThreads) {
        float v;
        if (from_slab)
          v = i < S->d3 ? stage[S->seg[SEG_BN_RM].pad_off + i]
                        : (i < 2 * S->d3 ? stage[S->seg[SEG_BN_RV].pad_off + i - S->d3] : stage[S->seg[SEG_BN_NBT].pad_off]);
        else
          v = S->wb_tmp[(size_t)p * S->nbuf + i];
        S->wb[(size_t)p * S->nbuf + i] = v;
      }
      if (threadIdx.x == 0) {
        S->ch.eta[p] = xt_at(S, L, e, sr, 0);          // MAIN:603
        S->ch.last_sse[p] = xt_at(S, L, e, sr, 3);
        S->ch.alias[p] = 0;                            // MAIN:602: w = dictfromlist(...) is a detached dict
      }
    }
    if (sr == rung) continue;
    // only trainable entries move into the live model; its BatchNorm buffers are reloaded from `w` next step
    float* dbase = S->theta + (size_t)p * S->Pp;
    for (int i = threadIdx.x; i < kAdamChunk / 4; i += kThreads) {
      const int pidx = it * kAdamChunk + i * 4;
      int sg = 0;
#pragma unroll
      for (int k = 1; k < kNumSegs; ++k) sg += (pidx >= seg_off[k]) ? 1 : 0;
      if (!seg_train[sg]) continue;
      *reinterpret_cast<float4*>(dbase + pidx) = *reinterpret_cast<const float4*>(stage + pidx);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The persistent program kernel
// ------------------------------------------------------------------------------------------------
template <bool kSimtGemm>
__device__ void run_phase(const DevState* S, int pi, int c_begin, int c_end, int force_swap, float* smem,
                          double* red) {
  const PhaseDesc& ph = S->phase[pi];
  if (ph.kind >= PH_SWAP_DECIDE && ph.kind <= PH_SWAP_SCATTER && !force_swap) {
    // inside bae_run: exchange only after every swap_interval-th iteration (MAIN:594); all chains share the step count
    const int st = S->ch.step[0];
    if (st == 0 || (st % S->swap_interval) != 0) return;
  }
  switch (ph.kind) {
    case PH_PROLOGUE: run_prologue(S, c_begin, c_end); break;
    case PH_GEMM:
      if constexpr (kSimtGemm) run_gemm_phase(S, ph, c_begin, c_end, smem, red);
      break;
    case PH_BN_FWD:
      if (S->d3 >= kBnCols) run_bn_fwd_wide(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
      else run_bn_fwd(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
      break;
    case PH_BN_BWD:
      if (S->d3 >= kBnCols) run_bn_bwd_wide(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
      else run_bn_bwd(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
      break;
    case PH_ADAM1: run_adam(S, 1, c_begin, c_end, red); break;
    case PH_ADAM2: run_adam(S, 2, c_begin, c_end, red); break;
    case PH_EPILOGUE: run_epilogue(S, c_begin, c_end, red); break;
    case PH_SWAP_DECIDE: run_swap_decide(S); break;
    case PH_SWAP_GATHER: run_swap_gather(S); break;
    case PH_SWAP_SCATTER: run_swap_scatter(S); break;
    case PH_WAMAX: run_wamax(S, ph, c_begin, c_end, force_swap); break;
    default: break;
  }
}

// Cooperative launch: walks phases [ph_begin, ph_end) `n_rep` times with a grid barrier after each.
__global__ void __launch_bounds__(kThreads, 2)
bae_program_kernel(const DevState* __restrict__ S, int ph_begin, int ph_end, int n_rep, int c_begin, int c_end,
                   int force_swap) {
  __shared__ __align__(16) float smem[2 * BK * 128 * 2];   // 32 KB: A and B double buffers (128-wide tiles)
  __shared__ double red[16];
  load_seg_table(S);
  cg::grid_group grid = cg::this_grid();
  for (int rep = 0; rep < n_rep; ++rep) {
    for (int pi = ph_begin; pi < ph_end; ++pi) {
      if (S->phase[pi].kind == PH_WAMAX) continue;   // 3xFP16 split only (per-phase launches)
      run_phase<true>(S, pi, c_begin, c_end, force_swap, smem, red);
      grid.sync();
    }
  }
}

// Plain launch of ONE phase (no grid barrier): used for per-phase profiling and the parity probe.
__global__ void __launch_bounds__(kThreads, 2)
bae_phase_kernel(const DevState* __restrict__ S, int pi, int c_begin, int c_end, int force_swap) {
  __shared__ __align__(16) float smem[2 * BK * 128 * 2];
  __shared__ double red[16];
  load_seg_table(S);
  run_phase<true>(S, pi, c_begin, c_end, force_swap, smem, red);
}

// The three exchange phases in ONE CTA (few chains, small parameter vectors: the narrow-width step kernel's companion) -
// no cooperative launch, no grid barriers.
__global__ void __launch_bounds__(kThreads, 1) bae_swap_cta_kernel(const DevState* __restrict__ S, int force_swap) {
  load_seg_table(S);
  if (!force_swap) {
    const int st = S->ch.step[0];
    if (st == 0 || (st % S->swap_interval) != 0) return;
  }
  run_swap_decide(S);
  __threadfence_block();
  __syncthreads();
  run_swap_gather(S);
  __threadfence_block();
  __syncthreads();
  run_swap_scatter(S);
}

// Non-GEMM phases of the tcgen05 path when every phase is its own launch (the default there): plain 256-thread CTAs,
// several per SM and with an L1 behind them, instead of the two epilogue warpgroups of the 1-CTA-per-SM tile engine.
// One kernel per phase family, so that each gets its own register allocation.
enum AuxGroup : int { AUX_ADAM = 0, AUX_BN = 1, AUX_MISC = 2 };
template <int kGroup>
__global__ void __launch_bounds__(kThreads, kGroup == AUX_ADAM ? kAuxAdamCtas : kGroup == AUX_BN ? kAuxBnCtas : 2)
bae_aux_kernel(const DevState* __restrict__ S, int pi, int c_begin, int c_end, int force_swap) {
  __shared__ __align__(16) float smem[1024];   // BatchNorm column reductions (512 doubles)
  __shared__ double red[16];
  load_seg_table(S);
  const PhaseDesc& ph = S->phase[pi];
  if constexpr (kGroup == AUX_ADAM) {
    run_adam(S, ph.kind == PH_ADAM1 ? 1 : 2, c_begin, c_end, red);
  } else if constexpr (kGroup == AUX_BN) {
    if (ph.kind == PH_BN_FWD) run_bn_fwd_wide(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
    else run_bn_bwd_wide(S, ph, c_begin, c_end, reinterpret_cast<double*>(smem));
  } else {
    run_phase<false>(S, pi, c_begin, c_end, force_swap, smem, red);
  }
}

// ---- tensor-core variants: GEMM phases go through the tcgen05 tile engine (1 CTA of 512 threads per SM).
// Warpgroups 1 and 2 (threads 128..383, 192 registers each) run the GEMM epilogues and every non-GEMM phase;
// warpgroups 0 and 3 (64 registers) hold the TMA producer, the MMA issuer and the hi/lo converters.
template <int kEpiA, int kEpiB>
__device__ __forceinline__ void run_phase_tc_back(const DevState* S, int pi, int c_begin, int c_end, int force_swap,
                                                  tc::Smem& tsm, tc::Ring& ring, float* smem_small, double* red) {
  const PhaseDesc& ph = S->phase[pi];
  if (ph.kind == PH_GEMM)
    tc::gemm_phase_epilogue<kEpiA, kEpiB>(S, ph, c_begin, c_end, tsm, ring);
  else if (ph.kind == PH_ADAM1 || ph.kind == PH_ADAM2)
    run_adam_tc(S, ph.kind == PH_ADAM1 ? 1 : 2, c_begin, c_end, red, tsm);
  else
    run_phase<false>(S, pi, c_begin, c_end, force_swap, smem_small, red);
  // generic-proxy stores of this phase (activations, theta, Y, dZ) are read by TMA (async proxy) in later phases
  tc::fence_proxy_async();
}

template <int kMode, bool kAK>
__device__ __forceinline__ void run_phase_tc_front(const DevState* S, int pi, int c_begin, int c_end, tc::Smem& tsm,
                                                   tc::Ring& ring) {
  const PhaseDesc& ph = S->phase[pi];
  if (ph.kind == PH_GEMM) tc::gemm_phase_front<kMode, kAK>(S, ph, c_begin, c_end, tsm, ring);
}

__global__ void __launch_bounds__(tc::kThreadsTc, 1)
bae_program_kernel_tc(const DevState* __restrict__ S, int ph_begin, int ph_end, int n_rep, int c_begin, int c_end,
                      int force_swap) {
  extern __shared__ uint8_t dyn_smem[];
  __shared__ double red[16];
  tc::Smem tsm;
  tc::Ring ring;
  tc::setup(tsm, dyn_smem);
  // BatchNorm column reductions (512 doubles) of the non-GEMM phases: the epilogue staging tiles, idle outside GEMM phases
  float* smem_small = tsm.epi_stage();
  load_seg_table(S);
  cg::grid_group grid = cg::this_grid();
  const int wg = threadIdx.x >> 7;
  if (wg == 1 || wg == 2) {
    tc::regs_inc<tc::kRegsEpiPersist>();
    for (int rep = 0; rep < n_rep; ++rep) {
      for (int pi = ph_begin; pi < ph_end; ++pi) {
        if (S->phase[pi].kind == PH_WAMAX) continue;   // 3xFP16 split only (per-phase launches)
        run_phase_tc_back<-1, -1>(S, pi, c_begin, c_end, force_swap, tsm, ring, smem_small, red);
        grid.sync();
        if (S->dbg && blockIdx.x == 0 && threadIdx.x == 128 && rep == 1) {   // debug: phase end times of the 2nd iteration
          long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          S->dbg[3584 + pi - ph_begin] = t;
        }
      }
    }
  } else {
    tc::regs_dec<tc::kRegsLeanPersist>();
    for (int rep = 0; rep < n_rep; ++rep) {
      for (int pi = ph_begin; pi < ph_end; ++pi) {
        if (S->phase[pi].kind == PH_WAMAX) continue;
        run_phase_tc_front<tc::FM_NONE, false>(S, pi, c_begin, c_end, tsm, ring);   // (3xTF32 only: no operand images)
        grid.sync();
      }
    }
  }
  tc::teardown(tsm);
}

// One GEMM phase as its own launch (the production path on tcgen05), specialised by the epilogue kinds of its GEMMs and by
// what they do with operand images (tc::FrontMode).
template <int kEpiA, int kEpiB, int kMode>
__global__ void __launch_bounds__(tc::kThreadsTc, 1)
bae_phase_kernel_tc(const DevState* __restrict__ S, int pi, int c_begin, int c_end, int force_swap) {
  extern __shared__ uint8_t dyn_smem[];
  tc::Smem tsm;
  tc::Ring ring;
  const bool out = kMode == tc::FM_B_OUT || (kMode == tc::FM_GENERIC && tc::phase_has_out(S, S->phase[pi]));
  const bool pre = kMode == tc::FM_GENERIC ? S->preconv != 0 : kMode != tc::FM_NONE;
  tc::setup(tsm, dyn_smem, S->f16 != 0, tc::kConvGroupWarps + (pre ? 1 : 0), out ? 2 : 1);
  const int wg = threadIdx.x >> 7;
  if (wg == 1 || wg == 2) {
    tc::regs_inc<tc::kRegsEpi>();
    tc::gemm_phase_epilogue<kEpiA, kEpiB>(S, S->phase[pi], c_begin, c_end, tsm, ring);
  } else {
    tc::regs_dec<tc::kRegsLean>();
    constexpr bool kAK = kMode != tc::FM_GENERIC && kMode != tc::FM_NONE && kEpiB < 0 && (kEpiA == EPI_SIGMOID || kEpiA == EPI_BIAS || kEpiA == EPI_LOSS);   // a forward layer
    run_phase_tc_front<kMode, kAK>(S, pi, c_begin, c_end, tsm, ring);
  }
  tc::teardown(tsm);
}

// Parity probe: fold the K-split gradient copies of one chain into copy 0.
__global__ void bae_grad_reduce_kernel(const DevState* __restrict__ S, int chain) {
  float* gr = S->grad + (size_t)chain * S->Pp;
  const size_t stride = (size_t)S->C_alloc * S->Pp;
  for (int sg = 0; sg < kNumSegs; ++sg) {
    const Seg& sd = S->seg[sg];
    const int ks = S->seg_ksplit[sg];
    if (ks <= 1) continue;
    const int n = sd.rows * sd.ld;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      float a = gr[sd.pad_off + e];
      for (int s2 = 1; s2 < ks; ++s2) a += gr[sd.pad_off + e + s2 * stride];
      gr[sd.pad_off + e] = a;
    }
  }
}

// value -> (hi, lo) TF32 split of n floats
__global__ void bae_split_kernel(const float* __restrict__ src, float* hi, float* lo, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float h, l;
    tc::split_tf32(src[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// Writes the proposal noise of (chain, step) in REFERENCE order, plus the three step scalars.
__global__ void bae_debug_draws_kernel(const DevState* __restrict__ S, int chain, int step, float* noise,
                                       double* scalars) {
  const int cg_id = S->ch.gid[chain];
  for (int sg = 0; sg < kNumSegs; ++sg) {
    const Seg& sd = S->seg[sg];
    const int n = sd.rows * sd.cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const int r = e / sd.cols, c = e % sd.cols;
      const int local = r * sd.ld + c;
      float nz[4];
      noise_quad(S, cg_id, step, sg, (unsigned)(local >> 2), nz);
      noise[sd.ref_off + e] = nz[local & 3] * S->step_w;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double lx, en, u;
    step_scalars(S, cg_id, step, lx, en, u);
    scalars[0] = lx;
    scalars[1] = en;
    scalars[2] = u;
  }
}

__global__ void bae_debug_swap_u_kernel(const DevState* __restrict__ S, int ensemble_global, int round, int pair,
                                        double* out) {
  out[0] = swap_uniform(S, ensemble_global, round, pair);
}

// ---- posterior-predictive extraction: dense [rows][d] copies of one chain's bottleneck features / reconstruction ----
// mode 0: out = src (features: Z, pitch ld);  mode 1: out = src / scale + X (the loss epilogue stored (o - x) * scale)
__global__ void bae_extract_kernel(const float* __restrict__ src, int ld, const float* __restrict__ X, int ldx, int rows, int d,
                                   float inv_scale, int mode, float* __restrict__ out) {
  const long long n = (long long)rows * d;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / d), c = (int)(e % d);
    const float v = src[(size_t)r * ld + c];
    out[e] = mode ? fmaf(v, inv_scale, X[(size_t)r * ldx + c]) : v;
  }
}

// ---- per-chain scalars <-> one staged struct (checkpoint / parity injection: one copy instead of 15) ----------
__global__ void bae_scalars_kernel(const DevState* __restrict__ S, int chain, bae_chain_scalars* st, int put) {
  const ChainArrays& c = S->ch;
  if (put) {
    c.eta[chain] = st->eta; c.lik[chain] = st->likelihood; c.prior_cur[chain] = st->prior_current;
    c.last_sse[chain] = st->last_sse_train; c.mse_tr[chain] = st->mse_train; c.mse_te[chain] = st->mse_test;
    c.T[chain] = st->temperature; c.adapt[chain] = st->adapttemp; c.adam_t[chain] = st->adam_t;
    c.alias[chain] = st->w_is_alias; c.init_count[chain] = st->init_count; c.nacc[chain] = st->num_accepted;
    c.nlang[chain] = st->langevin_count; c.step[chain] = st->step; c.prop_sse[chain] = st->prop_sse_train;
  } else {
    st->eta = c.eta[chain]; st->likelihood = c.lik[chain]; st->prior_current = c.prior_cur[chain];
    st->last_sse_train = c.last_sse[chain]; st->mse_train = c.mse_tr[chain]; st->mse_test = c.mse_te[chain];
    st->temperature = c.T[chain]; st->adapttemp = c.adapt[chain]; st->adam_t = c.adam_t[chain];
    st->w_is_alias = c.alias[chain]; st->init_count = c.init_count[chain]; st->num_accepted = c.nacc[chain];
    st->langevin_count = c.nlang[chain]; st->step = c.step[chain]; st->prop_sse_train = c.prop_sse[chain];
  }
}

// ---- reference-order <-> padded-order packing -----------------------------------------------------
__global__ void bae_pack_kernel(const DevState* __restrict__ S, const float* ref_vec, float* pad_vec) {
  for (int sg = 0; sg < kNumSegs; ++sg) {
    const Seg& sd = S->seg[sg];
    const int n = sd.rows * sd.cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const int r = e / sd.cols, c = e % sd.cols;
      pad_vec[sd.pad_off + r * sd.ld + c] = ref_vec[sd.ref_off + e];
    }
  }
}

__global__ void bae_unpack_kernel(const DevState* __restrict__ S, const float* pad_vec, float* ref_vec) {
  for (int sg = 0; sg < kNumSegs; ++sg) {
    const Seg& sd = S->seg[sg];
    const int n = sd.rows * sd.cols;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
      const int r = e / sd.cols, c = e % sd.cols;
      ref_vec[sd.ref_off + e] = pad_vec[sd.pad_off + r * sd.ld + c];
    }
  }
}

}  // namespace bae

// ---- launch wrappers used by bae_api.cu ----------------------------------------------------------
namespace bae {

// The tcgen05 kernels run as CTA pairs: cluster dimension 2, cooperative for the persistent kernel.
static cudaError_t launch_tc(const void* fn, int grid, void** args, bool cooperative, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::kThreadsTc); cfg.dynamicSmemBytes = tc::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = cooperative ? 2 : 1;
  return cudaLaunchKernelExC(&cfg, fn, args);
}

// The instantiated phase kernels: epilogue kinds (sorted, epi_b < 0: one kind) x front mode; null: not instantiated
static const void* phase_kernel_fn(int epi_a, int epi_b, int mode) {
#define BAE_PK(A, B, M) if (epi_a == (A) && epi_b == (B) && mode == (M)) return (const void*)bae_phase_kernel_tc<A, B, M>;
  // forward layers: generic; B image (test matrix pass); both images (first layer); B image + activation image out
  BAE_PK(EPI_SIGMOID, -1, tc::FM_NONE) BAE_PK(EPI_BIAS, -1, tc::FM_NONE) BAE_PK(EPI_LOSS, -1, tc::FM_NONE) BAE_PK(EPI_GRADW, -1, tc::FM_NONE)
  BAE_PK(EPI_DSIG, -1, tc::FM_NONE) BAE_PK(EPI_DSIG, EPI_GRADW, tc::FM_NONE) BAE_PK(EPI_PLAIN, EPI_GRADW, tc::FM_NONE)
  BAE_PK(EPI_SIGMOID, -1, tc::FM_GENERIC) BAE_PK(EPI_SIGMOID, -1, tc::FM_B) BAE_PK(EPI_SIGMOID, -1, tc::FM_AB) BAE_PK(EPI_SIGMOID, -1, tc::FM_B_OUT)
  BAE_PK(EPI_BIAS, -1, tc::FM_GENERIC) BAE_PK(EPI_BIAS, -1, tc::FM_B) BAE_PK(EPI_BIAS, -1, tc::FM_B_OUT)
  BAE_PK(EPI_LOSS, -1, tc::FM_GENERIC) BAE_PK(EPI_LOSS, -1, tc::FM_B) BAE_PK(EPI_LOSS, -1, tc::FM_B_OUT)
  // backward phases: generic; every B operand an image
  BAE_PK(EPI_GRADW, -1, tc::FM_GENERIC) BAE_PK(EPI_GRADW, -1, tc::FM_B)
  BAE_PK(EPI_DSIG, -1, tc::FM_GENERIC) BAE_PK(EPI_DSIG, -1, tc::FM_B)
  BAE_PK(EPI_DSIG, EPI_GRADW, tc::FM_GENERIC) BAE_PK(EPI_DSIG, EPI_GRADW, tc::FM_B)
  BAE_PK(EPI_PLAIN, EPI_GRADW, tc::FM_GENERIC) BAE_PK(EPI_PLAIN, EPI_GRADW, tc::FM_B)
#undef BAE_PK
  return nullptr;
}

int max_coop_blocks_per_sm(int use_tc) {
  int nb = 0;
  if (use_tc) {
    cudaFuncSetAttribute(bae_program_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    for (int ea = 0; ea <= EPI_GRADW; ++ea)
      for (int eb = -1; eb <= EPI_GRADW; ++eb)
        for (int m = 0; m < 5; ++m)
          if (const void* fn = phase_kernel_fn(ea, eb, m)) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bae_program_kernel_tc, tc::kThreadsTc, tc::kSmemBytes);
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bae_program_kernel, kThreads, 0);
  }
  return nb;
}

// CTAs of the persistent tcgen05 kernel: 2 x the number of co-resident CTA pairs.
int max_coop_grid_tc() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(tc::kThreadsTc); cfg.dynamicSmemBytes = tc::kSmemBytes;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, bae_program_kernel_tc, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return 2 * nclusters;
}

cudaError_t launch_program(const DevState* dS, int use_tc, int grid, int ph_begin, int ph_end, int n_rep, int c_begin,
                           int c_end, int force_swap, cudaStream_t st) {
  void* args[] = {(void*)&dS, (void*)&ph_begin, (void*)&ph_end, (void*)&n_rep, (void*)&c_begin, (void*)&c_end,
                  (void*)&force_swap};
  if (use_tc) return launch_tc((const void*)bae_program_kernel_tc, grid, args, true, st);
  return cudaLaunchCooperativeKernel((void*)bae_program_kernel, dim3(grid), dim3(kThreads), args, 0, st);
}

// epi_a / epi_b: epilogue kinds of the phase's GEMMs (host copy of the program; epi_b < 0: one GEMM)
cudaError_t launch_phase(const DevState* dS, int use_tc, int grid, int pi, int c_begin, int c_end, int force_swap,
                         int epi_a, int epi_b, int front_mode, cudaStream_t st) {
  if (use_tc) {
    void* args[] = {(void*)&dS, (void*)&pi, (void*)&c_begin, (void*)&c_end, (void*)&force_swap};
    if (epi_b >= 0 && epi_b < epi_a) { const int t = epi_a; epi_a = epi_b; epi_b = t; }
    if (epi_b == epi_a) epi_b = -1;
    const void* fn = phase_kernel_fn(epi_a, epi_b, front_mode);
    if (!fn) fn = phase_kernel_fn(epi_a, epi_b, tc::FM_GENERIC);
    if (!fn) return cudaErrorInvalidDeviceFunction;   // a GEMM-kind combination the step program does not contain
    return launch_tc(fn, grid, args, false, st);
  }
  bae_phase_kernel<<<grid, kThreads, 0, st>>>(dS, pi, c_begin, c_end, force_swap);
  return cudaGetLastError();
}

static int aux_group(int kind) {
  return (kind == PH_ADAM1 || kind == PH_ADAM2) ? AUX_ADAM : (kind == PH_BN_FWD || kind == PH_BN_BWD) ? AUX_BN : AUX_MISC;
}

// CTAs per SM of the aux kernel that runs phases of `kind`
int aux_blocks_per_sm(int kind) {
  int nb = 0;
  switch (aux_group(kind)) {
    case AUX_ADAM: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bae_aux_kernel<AUX_ADAM>, kThreads, 0); break;
    case AUX_BN: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bae_aux_kernel<AUX_BN>, kThreads, 0); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, bae_aux_kernel<AUX_MISC>, kThreads, 0); break;
  }
  return nb;
}

// `kind` is the phase's kind (host copy of the program); the tcgen05 path has d3 >= kBnCols, i.e. the wide BatchNorm code
cudaError_t launch_aux(const DevState* dS, int kind, int grid, int pi, int c_begin, int c_end, int force_swap,
                       cudaStream_t st) {
  switch (aux_group(kind)) {
    case AUX_ADAM: bae_aux_kernel<AUX_ADAM><<<grid, kThreads, 0, st>>>(dS, pi, c_begin, c_end, force_swap); break;
    case AUX_BN: bae_aux_kernel<AUX_BN><<<grid, kThreads, 0, st>>>(dS, pi, c_begin, c_end, force_swap); break;
    default: bae_aux_kernel<AUX_MISC><<<grid, kThreads, 0, st>>>(dS, pi, c_begin, c_end, force_swap); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_amax_zero(const DevState* dS, unsigned mask, int c_begin, int c_end, int gated, cudaStream_t st) {
  bae_amax_zero_kernel<<<1, 256, 0, st>>>(dS, mask, c_begin, c_end, gated);
  return cudaGetLastError();
}

cudaError_t launch_wconv(const DevState* dS, int grid, int c_begin, int c_end, int gated, cudaStream_t st) {
  bae_wconv_kernel<<<grid, kThreads, 0, st>>>(dS, c_begin, c_end, gated);
  return cudaGetLastError();
}

cudaError_t launch_ximg(const float* src, int ld, int R, int ccK, int ccM, int e, unsigned char* imgK, int kbsK, unsigned char* imgM,
                        int kbsM, cudaStream_t st) {
  bae_ximg_kernel<<<1184, kThreads, 0, st>>>(src, ld, R, ccK, ccM, e, imgK, kbsK, imgM, kbsM);
  return cudaGetLastError();
}

cudaError_t launch_himg_ones(unsigned char* img, long long stride, int kbs, int d, int rows, int nslots, cudaStream_t st) {
  bae_himg_ones_kernel<<<592, 256, 0, st>>>(img, stride, kbs, d, rows, nslots);
  return cudaGetLastError();
}

cudaError_t launch_swap_cta(const DevState* dS, int force_swap, cudaStream_t st) {
  bae_swap_cta_kernel<<<1, kThreads, 0, st>>>(dS, force_swap);
  return cudaGetLastError();
}

cudaError_t launch_grad_reduce(const DevState* dS, int chain, cudaStream_t st) {
  bae_grad_reduce_kernel<<<64, 256, 0, st>>>(dS, chain);
  return cudaGetLastError();
}

cudaError_t launch_split(const float* src, float* hi, float* lo, size_t n, cudaStream_t st) {
  bae_split_kernel<<<256, 256, 0, st>>>(src, hi, lo, n);
  return cudaGetLastError();
}

cudaError_t launch_debug_draws(const DevState* dS, int chain, int step, float* noise, double* scalars, cudaStream_t st) {
  bae_debug_draws_kernel<<<64, 256, 0, st>>>(dS, chain, step, noise, scalars);
  return cudaGetLastError();
}

cudaError_t launch_debug_swap_u(const DevState* dS, int ens, int round, int pair, double* out, cudaStream_t st) {
  bae_debug_swap_u_kernel<<<1, 1, 0, st>>>(dS, ens, round, pair, out);
  return cudaGetLastError();
}

cudaError_t launch_extract(const float* src, int ld, const float* X, int ldx, int rows, int d, float inv_scale, int mode,
                           float* out, cudaStream_t st) {
  bae_extract_kernel<<<592, 256, 0, st>>>(src, ld, X, ldx, rows, d, inv_scale, mode, out);
  return cudaGetLastError();
}

cudaError_t launch_xtable(const DevState* dS, cudaStream_t st) { bae_xtable_kernel<<<8, 256, 0, st>>>(dS); return cudaGetLastError(); }
cudaError_t launch_xsweep(const DevState* dS, cudaStream_t st) { bae_xsweep_kernel<<<4, 64, 0, st>>>(dS); return cudaGetLastError(); }
cudaError_t launch_xgather(const DevState* dS, int grid, cudaStream_t st) { bae_xgather_kernel<<<grid, kThreads, 0, st>>>(dS); return cudaGetLastError(); }
cudaError_t launch_xscatter(const DevState* dS, int grid, cudaStream_t st) { bae_xscatter_kernel<<<grid, kThreads, 0, st>>>(dS); return cudaGetLastError(); }

cudaError_t launch_scalars(const DevState* dS, int chain, bae_chain_scalars* staged, int put, cudaStream_t st) {
  bae_scalars_kernel<<<1, 1, 0, st>>>(dS, chain, staged, put);
  return cudaGetLastError();
}

cudaError_t launch_pack(const DevState* dS, const float* ref_vec, float* pad_vec, cudaStream_t st) {
  bae_pack_kernel<<<64, 256, 0, st>>>(dS, ref_vec, pad_vec);
  return cudaGetLastError();
}

cudaError_t launch_unpack(const DevState* dS, const float* pad_vec, float* ref_vec, cudaStream_t st) {
  bae_unpack_kernel<<<64, 256, 0, st>>>(dS, pad_vec, ref_vec);
  return cudaGetLastError();
}

}  // namespace bae
