// Narrow-width step kernel (sm_100a): the whole PT-Langevin iteration of MAIN:432-592 for networks whose layers are a few
// units wide (Swiss Roll: 3-10-5-2-5-10-3, MAIN:97-100), where a layer is not a dense contraction: FP32 CUDA cores, one
// data row per thread with all activations of the row in registers, warp-shuffle + shared-memory reductions inside a CTA and
// distributed-shared-memory reductions across a THREAD-BLOCK CLUSTER.
//
//   one cluster (1, 2, 4 or 8 CTAs of 256 threads) per chain; the data rows are cut into contiguous ranges, one per CTA
//   the chain's state (live model theta, Adam m / v, the detached-w buffers, every scalar) lives in shared memory,
//     REPLICATED in each CTA of the cluster: all CTAs execute the same Adam / Metropolis-Hastings arithmetic on the same
//     reduced sums (bit-identical), so nothing has to be broadcast; cluster rank 0 writes traces and the final state
//   `n_iter` iterations per launch (the exchange sweep between launches is the generic swap phase program)
//
// One forward / backward over this CTA's rows (a drift call, MAIN:208-226):
//   pass A  x -> h1 -> h2 -> z per row, z kept in shared memory; sum z and sum z^2 in FP64 (batch statistics)
//                                                                                           [1 cluster reduction]
//   pass B  z -> BatchNorm -> h4 -> h5 -> out, loss, deltas back to dy, in chunks of 256 rows: per-row values are staged
//           in shared memory and the decoder parameters are dot products over the staged rows (a thread = one delta column
//           x four activation columns x 64 rows: one scalar and one 16-byte shared-memory load per four FMAs);
//           sum dy, sum dy * zhat (BatchNorm backward and the gamma / beta gradients), SSE   [1 cluster reduction]
//   pass C  dz from the BatchNorm backward formula, deltas back to layer 1 (h1, h2 recomputed from x), encoder parameters
//           as dot products; the per-CTA partial gradient of all parameters                 [1 cluster reduction]
// The arithmetic mirrors the generic phases of bae_kernels.cu (run_prologue, run_bn_*, run_adam, run_epilogue) - same
// formulas, same Philox streams, same state layout in HBM - so every host entry point (init, probe, exchange, get / set
// state) works unchanged on chains advanced by this kernel.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "bae_common.cuh"

namespace cg = cooperative_groups;

namespace bae {
namespace narrow {

constexpr int kT = 256;
constexpr int kMaxP = 768;          // floats of the padded parameter vector that are in use (Swiss: 704)
constexpr int kMaxRows = 4096;      // rows per CTA whose z / dy are kept in shared memory
constexpr float kBnEps = 1e-5f, kBnMom = 0.1f, kB1 = 0.9f, kB2 = 0.999f, kAdamEps = 1e-8f;

__host__ __device__ constexpr int r4(int x) { return (x + 3) / 4 * 4; }

// same form as the tensor-core epilogue (bae_tc.cuh): two MUFU operations
__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct SegN { int pad_off, span, cols, ld, trainable; };

// chain scalars mirrored in shared memory (ChainArrays of bae_common.cuh)
struct ChainSm {
  double eta, lik, prior_cur, last_sse, prop_sse, mse_tr, mse_te, T, adapt;
  double lx, eta_noise, u;
  float ss1, sq1, ss2, sq2;
  int adam_t, alias, init_count, nacc, nlang, step, is_lang, gid, accept;
};

template <int D0, int D1, int D2, int D3>
struct Smem {
  static constexpr int NB = 2 * D3 + 1;
  // staged floats per row: three delta blocks and three [activation | 1] blocks, each padded to a multiple of 4 floats
  static constexpr int SWB = r4(D0) + r4(D1) + r4(D2) + r4(D1 + 1) + r4(D2 + 1) + r4(D3 + 1);
  static constexpr int SWC = r4(D3) + r4(D2) + r4(D1) + r4(D2) + r4(D1 + 1) + r4(D0 + 1);
  static constexpr int SW0 = SWB > SWC ? SWB : SWC;
  static constexpr int SW = (SW0 % 32 == 20) ? SW0 : SW0 + ((20 - SW0 % 32 + 32) % 32);   // pitch = 20 mod 32 floats: 16-byte row stores of a warp hit distinct banks
  float th[kMaxP], m[kMaxP], v[kMaxP], wc[kMaxP], grad[kMaxP];
  float gpart[2][kMaxP];   // this CTA's partial gradient, double-buffered; the peers read it through DSMEM
  float dotq[4][kMaxP];    // dot-product partials of the four 64-row groups of the staged chunks
  double dpart[2][8];      // this CTA's partial FP64 sums, double-buffered
  double dsum[8];          // cluster totals
  double red[64];          // block reduction scratch
  float wb[NB], wb_tmp[NB];
  float mu[3][D3], var[3][D3], inv[D3];
  SegN seg[kNumSegs];
  ChainSm c;
  float zs[kMaxRows * D3];
  float dys[kMaxRows * D3];
  alignas(16) float stage[kT * SW];
};

// ---- reductions ------------------------------------------------------------------------------------------------
// Sum of NV doubles over the 256 threads of the CTA, fixed order; result in every thread.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) red[w * NV + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kT / 32; ++i) t += red[i * NV + k];
    v[k] = t;
  }
}

// Cluster total of this CTA's `n` partial doubles (<= 8): partial -> own shared memory, cluster barrier, every CTA adds the
// partials of all ranks in rank order through distributed shared memory. `buf` alternates between consecutive reductions:
// a slot is rewritten two reductions later, after a barrier that every rank passes only when it has read the slot.
template <class SM>
__device__ __forceinline__ void cluster_sum(cg::cluster_group& cl, SM& sm, const double* part, int n, int& buf) {
  if (threadIdx.x == 0)
    for (int k = 0; k < n; ++k) sm.dpart[buf][k] = part[k];
  cl.sync();
  if (threadIdx.x < n) {
    double t = 0.0;
    const unsigned nr = cl.num_blocks();
    for (unsigned r = 0; r < nr; ++r) t += *cl.map_shared_rank(&sm.dpart[buf][threadIdx.x], r);
    sm.dsum[threadIdx.x] = t;
  }
  __syncthreads();
  buf ^= 1;
}

// ---- layer helpers (weights in shared memory, one row in registers) --------------------------------------------------
template <int NO, int NI, bool SIG>
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float (&in)[NI], float (&out)[NO]) {
  constexpr int LD = r4(NI);
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    float a = b[o];
#pragma unroll
    for (int i = 0; i < NI; ++i) a = fmaf(in[i], W[o * LD + i], a);
    out[o] = SIG ? sigm(a) : a;
  }
}
// din[i] = sum_o dout[o] * W[o][i]
template <int NO, int NI>
__device__ __forceinline__ void dense_t(const float* __restrict__ W, const float (&dout)[NO], float (&din)[NI]) {
  constexpr int LD = r4(NI);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    float a = 0.f;
#pragma unroll
    for (int o = 0; o < NO; ++o) a = fmaf(dout[o], W[o * LD + i], a);
    din[i] = a;
  }
}

// ---- Philox draws: identical to noise_quad / step_scalars / box_muller of bae_kernels.cu -------------------------------
__device__ __forceinline__ void box_muller(float u1, float u2, float& n0, float& n1) {
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}
__device__ __forceinline__ void noise_quad(unsigned k0, unsigned k1, int gid, int step, int seg, unsigned quad, float (&n)[4]) {
  u4 c;
  c.x = quad; c.y = (unsigned)seg | (STREAM_WEIGHT_NOISE << 8); c.z = (unsigned)step; c.w = (unsigned)gid;
  const u4 r = philox4x32_10(c, k0, k1);
  box_muller(u01(r.x), u01(r.y), n[0], n[1]);
  box_muller(u01(r.z), u01(r.w), n[2], n[3]);
}
__device__ __forceinline__ double loglik_d(double sse, double n, double tau_sq, double temp) {
  return (-(n * 0.5) * log(2.0 * CUDART_PI * tau_sq) - 0.5 * tau_sq * sse) / temp;   // MAIN:325-329 (tau_sq multiplies the SSE)
}

template <class SM>
__device__ __forceinline__ int find_seg(const SM& sm, int p) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < kNumSegs; ++i) s += (p >= sm.seg[i].pad_off) ? 1 : 0;
  return s;
}

// =====================================================================================================================
// 16-byte stores of one block of a staged row: the N values, a 1.0 when the block carries the ones column (bias gradients),
// zeros up to the next multiple of 4
template <int N, bool ONE>
__device__ __forceinline__ void store_block(float* __restrict__ st, const float (&v)[N]) {
  constexpr int W = r4(N + (ONE ? 1 : 0));
  float t[W];
#pragma unroll
  for (int i = 0; i < W; ++i) t[i] = i < N ? v[i] : ((ONE && i == N) ? 1.0f : 0.0f);
#pragma unroll
  for (int i = 0; i < W; i += 4) *reinterpret_cast<float4*>(st + i) = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
}

template <int D0, int D1, int D2, int D3>
struct Net {
  using SM = Smem<D0, D1, D2, D3>;
  static constexpr int SW = SM::SW;
  // staged row of pass B (decoder): dout | d5 | d4 | [h5 1] | [h4 1] | [y 1]      pass C (encoder): dz | d2 | d1 | h2 | [h1 1] | [x 1]
  static constexpr int B_DOUT = 0, B_D5 = r4(D0), B_D4 = B_D5 + r4(D1), B_H5 = B_D4 + r4(D2), B_H4 = B_H5 + r4(D1 + 1), B_Y = B_H4 + r4(D2 + 1);
  static constexpr int C_DZ = 0, C_D2 = r4(D3), C_D1 = C_D2 + r4(D2), C_H2 = C_D1 + r4(D1), C_H1 = C_H2 + r4(D2), C_X = C_H1 + r4(D1 + 1);
  __host__ __device__ static constexpr int q4(int n) { return (n + 3) / 4; }
  // A dot GROUP = one delta column (a) times one 4-float block of an [activation | 1] block (b): four gradient entries of one
  // weight row (the ones column yields the bias gradient). A thread takes one group over the 64 rows of one row group.
  static constexpr int GD0 = D0 * q4(D1 + 1), GD1 = D1 * q4(D2 + 1), GD2 = D2 * q4(D3 + 1), GD = GD0 + GD1 + GD2;   // W6|b6, W5|b5, W4|b4
  static constexpr int GE0 = D3 * q4(D2), GE1 = D2 * q4(D1 + 1), GE2 = D1 * q4(D0 + 1), GE = GE0 + GE1 + GE2;       // W3, W2|b2, W1|b1
  static_assert(GD <= 64 && GE <= 64, "64 groups x 4 row groups = 256 threads");
  static_assert(SM::SWB <= SW && SM::SWC <= SW && SW % 4 == 0, "stage width");

  struct Dot { int a, b, dst[4]; };
  __device__ static void fill(Dot& d, int g, int nb, int NI, int acol0, int bcol0, int w_off, int b_off) {
    const int o = g / nb, jb = g % nb;
    d.a = acol0 + o;
    d.b = bcol0 + jb * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = jb * 4 + j;
      d.dst[j] = i < NI ? w_off + o * r4(NI) + i : ((i == NI && b_off >= 0) ? b_off + o : -1);
    }
  }
  __device__ static Dot dec_dot(const SM& sm, int g) {
    Dot d{0, 0, {-1, -1, -1, -1}};
    if (g < GD0) fill(d, g, q4(D1 + 1), D1, B_DOUT, B_H5, sm.seg[SEG_D5_W].pad_off, sm.seg[SEG_D5_B].pad_off);
    else if (g < GD0 + GD1) fill(d, g - GD0, q4(D2 + 1), D2, B_D5, B_H4, sm.seg[SEG_D3_W].pad_off, sm.seg[SEG_D3_B].pad_off);
    else if (g < GD) fill(d, g - GD0 - GD1, q4(D3 + 1), D3, B_D4, B_Y, sm.seg[SEG_D1_W].pad_off, sm.seg[SEG_D1_B].pad_off);
    return d;
  }
  __device__ static Dot enc_dot(const SM& sm, int g) {
    Dot d{0, 0, {-1, -1, -1, -1}};
    if (g < GE0) fill(d, g, q4(D2), D2, C_DZ, C_H2, sm.seg[SEG_E4_W].pad_off, -1);   // encode.4.bias: gradient exactly 0 (BatchNorm removes it)
    else if (g < GE0 + GE1) fill(d, g - GE0, q4(D1 + 1), D1, C_D2, C_H1, sm.seg[SEG_E2_W].pad_off, sm.seg[SEG_E2_B].pad_off);
    else if (g < GE) fill(d, g - GE0 - GE1, q4(D0 + 1), D0, C_D1, C_X, sm.seg[SEG_E0_W].pad_off, sm.seg[SEG_E0_B].pad_off);
    return d;
  }
  // 64 staged rows of this thread's row group: acc[j] += a * b[j]
  __device__ static __forceinline__ void dot64(const float* __restrict__ stage, const Dot& d, float (&acc)[4]) {
    const float* a = stage + (threadIdx.x >> 6) * 64 * SW + d.a;
    const float* b = stage + (threadIdx.x >> 6) * 64 * SW + d.b;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float av = a[k * SW];
      const float4 bv = *reinterpret_cast<const float4*>(b + k * SW);
      acc[0] = fmaf(av, bv.x, acc[0]); acc[1] = fmaf(av, bv.y, acc[1]); acc[2] = fmaf(av, bv.z, acc[2]); acc[3] = fmaf(av, bv.w, acc[3]);
    }
  }

  __device__ static void encoder(const SM& sm, const float (&x)[D0], float (&h1)[D1], float (&h2)[D2], float (&z)[D3]) {
    dense<D1, D0, true>(sm.th + sm.seg[SEG_E0_W].pad_off, sm.th + sm.seg[SEG_E0_B].pad_off, x, h1);
    dense<D2, D1, true>(sm.th + sm.seg[SEG_E2_W].pad_off, sm.th + sm.seg[SEG_E2_B].pad_off, h1, h2);
    dense<D3, D2, false>(sm.th + sm.seg[SEG_E4_W].pad_off, sm.th + sm.seg[SEG_E4_B].pad_off, h2, z);
  }

  // Forward (and backward when `back`) of the rows [r0, r1) of X [n_rows x D0] (pitch ldx) at the weights in sm.th.
  // Leaves: batch statistics in sm.mu / sm.var [slot] and sm.inv; cluster total sm.dsum[0] = SSE; with `back` the full
  // gradient in sm.grad. `buf_d` / `buf_g` are the double-buffer indices of the cluster reductions.
  __device__ static void pass(cg::cluster_group& cl, SM& sm, const float* __restrict__ X, int ldx, int n_rows, int r0, int r1, int slot,
                              bool back, int& buf_d, int& buf_g) {
    const int tid = threadIdx.x;
    const int nr = r1 - r0;
    // ---- pass A: z of every row; batch statistics from sum z and sum z^2 in FP64
    {
      double sz[2 * D3];
#pragma unroll
      for (int c = 0; c < 2 * D3; ++c) sz[c] = 0.0;
      for (int r = tid; r < nr; r += kT) {
        float x[D0], h1[D1], h2[D2], z[D3];
#pragma unroll
        for (int i = 0; i < D0; ++i) x[i] = __ldg(&X[(size_t)(r0 + r) * ldx + i]);
        encoder(sm, x, h1, h2, z);
#pragma unroll
        for (int c = 0; c < D3; ++c) {
          sm.zs[r * D3 + c] = z[c];
          sz[c] += (double)z[c];
          sz[D3 + c] += (double)z[c] * (double)z[c];
        }
      }
      block_sum<2 * D3>(sz, sm.red);
      cluster_sum(cl, sm, sz, 2 * D3, buf_d);
      if (tid < D3) {
        const double mean = sm.dsum[tid] / n_rows;
        const float var = (float)fmax(sm.dsum[D3 + tid] / n_rows - mean * mean, 0.0);   // biased batch variance
        sm.mu[slot][tid] = (float)mean;
        sm.var[slot][tid] = var;
        sm.inv[tid] = 1.0f / sqrtf(var + kBnEps);
      }
      if (back)
        for (int i = tid; i < 4 * kMaxP; i += kT) (&sm.dotq[0][0])[i] = 0.f;
      __syncthreads();
    }
    const float* gam = sm.th + sm.seg[SEG_BN_W].pad_off;
    const float* bet = sm.th + sm.seg[SEG_BN_BIAS].pad_off;
    const float scale = 2.0f / ((float)n_rows * (float)D0);   // d MSELoss / d out (mean over all N*d0 elements, MAIN:150,223)
    const int q = tid >> 6, g = tid & 63;
    // ---- pass B: decoder forward, loss, decoder backward; decoder gradients as dot products over the staged rows
    double acc3[1 + 2 * D3];   // SSE, sum dy[c], sum dy[c] * zhat[c]
#pragma unroll
    for (int k = 0; k < 1 + 2 * D3; ++k) acc3[k] = 0.0;
    float accD[4] = {0.f, 0.f, 0.f, 0.f};
    const Dot dd = dec_dot(sm, g);
    for (int c0 = 0; c0 < nr; c0 += kT) {
      const int r = c0 + tid;
      float* st = sm.stage + tid * SW;
      if (r < nr) {
        float x[D0], zhat[D3], y[D3], h4[D2], h5[D1], out[D0];
#pragma unroll
        for (int i = 0; i < D0; ++i) x[i] = __ldg(&X[(size_t)(r0 + r) * ldx + i]);
#pragma unroll
        for (int c = 0; c < D3; ++c) {
          zhat[c] = (sm.zs[r * D3 + c] - sm.mu[slot][c]) * sm.inv[c];
          y[c] = fmaf(zhat[c], gam[c], bet[c]);
        }
        dense<D2, D3, true>(sm.th + sm.seg[SEG_D1_W].pad_off, sm.th + sm.seg[SEG_D1_B].pad_off, y, h4);
        dense<D1, D2, true>(sm.th + sm.seg[SEG_D3_W].pad_off, sm.th + sm.seg[SEG_D3_B].pad_off, h4, h5);
        dense<D0, D1, false>(sm.th + sm.seg[SEG_D5_W].pad_off, sm.th + sm.seg[SEG_D5_B].pad_off, h5, out);
        float dout[D0];
        float l_sse = 0.f;
#pragma unroll
        for (int o = 0; o < D0; ++o) {
          const float d = out[o] - x[o];
          l_sse = fmaf(d, d, l_sse);
          dout[o] = d * scale;
        }
        acc3[0] += (double)l_sse;
        if (back) {
          float d5[D1], d4[D2], dy[D3];
          dense_t<D0, D1>(sm.th + sm.seg[SEG_D5_W].pad_off, dout, d5);
#pragma unroll
          for (int i = 0; i < D1; ++i) d5[i] *= h5[i] * (1.0f - h5[i]);
          dense_t<D1, D2>(sm.th + sm.seg[SEG_D3_W].pad_off, d5, d4);
#pragma unroll
          for (int i = 0; i < D2; ++i) d4[i] *= h4[i] * (1.0f - h4[i]);
          dense_t<D2, D3>(sm.th + sm.seg[SEG_D1_W].pad_off, d4, dy);
#pragma unroll
          for (int c = 0; c < D3; ++c) {
            sm.dys[r * D3 + c] = dy[c];
            acc3[1 + c] += (double)dy[c];
            acc3[1 + D3 + c] += (double)dy[c] * (double)zhat[c];
          }
          store_block<D0, false>(st + B_DOUT, dout);
          store_block<D1, false>(st + B_D5, d5);
          store_block<D2, false>(st + B_D4, d4);
          store_block<D1, true>(st + B_H5, h5);
          store_block<D2, true>(st + B_H4, h4);
          store_block<D3, true>(st + B_Y, y);
        }
      } else if (back) {
#pragma unroll
        for (int i = 0; i < SM::SWB; i += 4) *reinterpret_cast<float4*>(st + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (back) {
        __syncthreads();
        if (g < GD) dot64(sm.stage, dd, accD);
        __syncthreads();
      }
    }
    block_sum<1 + 2 * D3>(acc3, sm.red);
    cluster_sum(cl, sm, acc3, 1 + 2 * D3, buf_d);
    if (!back) return;   // sm.dsum[0] = SSE
    const double sse_total = sm.dsum[0];
    float m1[D3], m2[D3];
#pragma unroll
    for (int c = 0; c < D3; ++c) {
      m1[c] = (float)(sm.dsum[1 + c] / n_rows) * gam[c];        // mean(dzhat), dzhat = dy * gamma
      m2[c] = (float)(sm.dsum[1 + D3 + c] / n_rows) * gam[c];   // mean(dzhat * zhat)
    }
    const float g_beta = tid < D3 ? (float)sm.dsum[1 + tid] : 0.f;          // d beta  = sum dy
    const float g_gamma = tid < D3 ? (float)sm.dsum[1 + D3 + tid] : 0.f;    // d gamma = sum dy * zhat
    // ---- pass C: BatchNorm backward, encoder backward (h1, h2 recomputed from x), encoder gradients as dot products
    float accE[4] = {0.f, 0.f, 0.f, 0.f};
    const Dot de = enc_dot(sm, g);
    for (int c0 = 0; c0 < nr; c0 += kT) {
      const int r = c0 + tid;
      float* st = sm.stage + tid * SW;
      if (r < nr) {
        float x[D0], h1[D1], h2[D2], z[D3], dz[D3], d2[D2], d1[D1];
#pragma unroll
        for (int i = 0; i < D0; ++i) x[i] = __ldg(&X[(size_t)(r0 + r) * ldx + i]);
        encoder(sm, x, h1, h2, z);
#pragma unroll
        for (int c = 0; c < D3; ++c) {
          const float zhat = (sm.zs[r * D3 + c] - sm.mu[slot][c]) * sm.inv[c];
          dz[c] = sm.inv[c] * (sm.dys[r * D3 + c] * gam[c] - m1[c] - zhat * m2[c]);
        }
        dense_t<D3, D2>(sm.th + sm.seg[SEG_E4_W].pad_off, dz, d2);
#pragma unroll
        for (int i = 0; i < D2; ++i) d2[i] *= h2[i] * (1.0f - h2[i]);
        dense_t<D2, D1>(sm.th + sm.seg[SEG_E2_W].pad_off, d2, d1);
#pragma unroll
        for (int i = 0; i < D1; ++i) d1[i] *= h1[i] * (1.0f - h1[i]);
        store_block<D3, false>(st + C_DZ, dz);
        store_block<D2, false>(st + C_D2, d2);
        store_block<D1, false>(st + C_D1, d1);
        store_block<D2, false>(st + C_H2, h2);
        store_block<D1, true>(st + C_H1, h1);
        store_block<D0, true>(st + C_X, x);
      } else {
#pragma unroll
        for (int i = 0; i < SM::SWC; i += 4) *reinterpret_cast<float4*>(st + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      if (g < GE) dot64(sm.stage, de, accE);
      __syncthreads();
    }
    // ---- partial gradient of this CTA: the four row groups in fixed order, then the cluster total in rank order
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (g < GD && dd.dst[j] >= 0) sm.dotq[q][dd.dst[j]] = accD[j];
      if (g < GE && de.dst[j] >= 0) sm.dotq[q][de.dst[j]] = accE[j];
    }
    __syncthreads();
    float* gp = sm.gpart[buf_g];
    for (int i = tid; i < kMaxP; i += kT) gp[i] = ((sm.dotq[0][i] + sm.dotq[1][i]) + sm.dotq[2][i]) + sm.dotq[3][i];
    cl.sync();
    {
      const unsigned nrk = cl.num_blocks();
      for (int i = tid; i < kMaxP; i += kT) {
        float t = 0.f;
        for (unsigned rk = 0; rk < nrk; ++rk) t += *cl.map_shared_rank(&gp[i], rk);
        sm.grad[i] = t;
      }
    }
    __syncthreads();
    if (tid < D3) {
      sm.grad[sm.seg[SEG_BN_BIAS].pad_off + tid] = g_beta;
      sm.grad[sm.seg[SEG_BN_W].pad_off + tid] = g_gamma;
    }
    if (tid == 0) sm.dsum[0] = sse_total;
    __syncthreads();
    buf_g ^= 1;
  }

  // Adam drift + proposal noise (which = 1, MAIN:473-474 / random walk MAIN:492) and the reverse drift (which = 2,
  // MAIN:475-477), with the sums the MH ratio needs: out[0] = sum noise^2, out[1] = sum proposal^2, out[2] = sum (w - w_prop_gd)^2.
  // Same arithmetic as run_adam (bae_kernels.cu) on the shared-memory copy of the state; every CTA of the cluster does all of it.
  __device__ static void adam(SM& sm, int which, int p_used, unsigned k0, unsigned k1, float step_w, double (&out)[3]) {
    const bool lang = sm.c.is_lang != 0, alias = sm.c.alias != 0;
    const float ss = which == 1 ? sm.c.ss1 : sm.c.ss2, sq = which == 1 ? sm.c.sq1 : sm.c.sq2;
    const float inv_sq = 1.0f / sq;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (!(which == 2 && !lang))
      for (int p = threadIdx.x * 4; p < p_used; p += kT * 4) {
        const int sg = find_seg(sm, p);
        const SegN sd = sm.seg[sg];
        if (!sd.trainable) continue;
        const int local = p - sd.pad_off;
        if (local >= sd.span) continue;
        const int col = local % sd.ld;
        const int nval = min(4, sd.cols - col);
        if (nval <= 0) continue;
        float t[4], g[4], m[4], v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { t[j] = sm.th[p + j]; g[j] = sm.grad[p + j]; m[j] = sm.m[p + j]; v[j] = sm.v[p + j]; }
        if (which == 1) {
          float o[4];
          if (lang) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
              v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
              const float denom = fmaf(sqrtf(v[j]), inv_sq, kAdamEps);
              o[j] = t[j] - ss * __fdividef(m[j], denom);
              sm.m[p + j] = m[j];
              sm.v[p + j] = v[j];
              if (!alias) sm.wc[p + j] = t[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = t[j];
          }
          float nz[4];
          noise_quad(k0, k1, sm.c.gid, sm.c.step, sg, (unsigned)(local >> 2), nz);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nval) {
              const float nj = nz[j] * step_w;
              o[j] = o[j] + nj;
              a0 = fmaf(nj, nj, a0);
              a1 = fmaf(o[j], o[j], a1);
            } else {
              o[j] = 0.f;
            }
            sm.th[p + j] = o[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[j] = m[j] + (g[j] - m[j]) * (1.0f - kB1);
            v[j] = v[j] * kB2 + (1.0f - kB2) * g[j] * g[j];
            const float denom = fmaf(sqrtf(v[j]), inv_sq, kAdamEps);
            const float wpg = t[j] - ss * __fdividef(m[j], denom);   // w_prop_gd
            if (!alias && j < nval) {
              const float dl = sm.wc[p + j] - wpg;
              a2 = fmaf(dl, dl, a2);
            }
            sm.m[p + j] = m[j];
            sm.v[p + j] = v[j];
          }
        }
      }
    double s[3] = {(double)a0, (double)a1, (double)a2};
    block_sum<3>(s, sm.red);
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2];
  }
};

// ---------------------------------------------------------------------------------------------------------------------
template <int D0, int D1, int D2, int D3>
__global__ void __launch_bounds__(kT, 1) bae_narrow_kernel(const DevState* __restrict__ S, int n_iter, int p_used) {
  using N = Net<D0, D1, D2, D3>;
  using SM = typename N::SM;
  extern __shared__ __align__(16) uint8_t dyn[];
  SM& sm = *reinterpret_cast<SM*>(dyn);
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  const unsigned CL = cl.num_blocks(), rank = cl.block_rank();
  const int chain = blockIdx.x / CL;
  const ChainArrays& ch = S->ch;
  const unsigned k0 = S->seed_lo, k1 = S->seed_hi;
  const int NB = SM::NB;
  // ---- state of the chain -> shared memory (every CTA of the cluster holds a copy)
  if (tid < kNumSegs) {
    const Seg& g = S->seg[tid];
    sm.seg[tid] = SegN{g.pad_off, g.rows * g.ld, g.cols, g.ld, g.trainable};
  }
  for (int i = tid; i < kMaxP; i += kT) {
    const bool in = i < p_used;
    sm.th[i] = in ? S->theta[(size_t)chain * S->Pp + i] : 0.f;
    sm.m[i] = in ? S->m[(size_t)chain * S->Pp + i] : 0.f;
    sm.v[i] = in ? S->v[(size_t)chain * S->Pp + i] : 0.f;
    sm.wc[i] = 0.f;
    sm.grad[i] = 0.f;
  }
  if (tid < NB) sm.wb[tid] = S->wb[(size_t)chain * S->nbuf + tid];
  if (tid == 0) {
    ChainSm& c = sm.c;
    c.eta = ch.eta[chain]; c.lik = ch.lik[chain]; c.prior_cur = ch.prior_cur[chain]; c.last_sse = ch.last_sse[chain];
    c.prop_sse = ch.prop_sse[chain]; c.mse_tr = ch.mse_tr[chain]; c.mse_te = ch.mse_te[chain]; c.T = ch.T[chain]; c.adapt = ch.adapt[chain];
    c.adam_t = ch.adam_t[chain]; c.alias = ch.alias[chain]; c.init_count = ch.init_count[chain]; c.nacc = ch.nacc[chain];
    c.nlang = ch.nlang[chain]; c.step = ch.step[chain]; c.gid = ch.gid[chain];
  }
  __syncthreads();
  cl.sync();   // every CTA's shared memory exists before any peer maps it
  const int n_train = S->n_train, n_test = S->n_test, ldx = S->ld0;
  const int per_tr = (n_train + (int)CL - 1) / (int)CL, per_te = (n_test + (int)CL - 1) / (int)CL;
  const int tr0 = min(n_train, (int)rank * per_tr), tr1 = min(n_train, tr0 + per_tr);
  const int te0 = min(n_test, (int)rank * per_te), te1 = min(n_test, te0 + per_te);
  int buf_d = 0, buf_g = 0;
  const float step_w = S->step_w;
  const int off_rm = sm.seg[SEG_BN_RM].pad_off, off_rv = sm.seg[SEG_BN_RV].pad_off, off_nbt = sm.seg[SEG_BN_NBT].pad_off;

  for (int it = 0; it < n_iter; ++it) {
    // ---- prologue: branch select, tempered -> canonical switch, draws   (run_prologue; MAIN:435-452, 500, 526)
    if (tid == 0) {
      ChainSm& c = sm.c;
      const int i = c.step;
      if (i < S->pt_step) c.adapt = c.T;
      if (i == S->pt_step && c.init_count == 0) {
        c.adapt = 1.0;
        c.lik = loglik_d(c.prop_sse, (double)n_train * D0, 1.0, 1.0);
        c.init_count = 1;
      }
      u4 ctr;
      ctr.x = 0u; ctr.y = 0xFFu | (STREAM_STEP_SCALARS << 8); ctr.z = (unsigned)i; ctr.w = (unsigned)c.gid;
      const u4 r = philox4x32_10(ctr, k0, k1);
      c.lx = (double)u01(r.x);
      float n0, n1;
      box_muller(u01(r.y), u01(r.z), n0, n1);
      c.eta_noise = (double)n0 * (double)S->step_eta;
      c.u = (double)u01(r.w);
      c.is_lang = (S->use_langevin && c.lx < (double)S->l_prob) ? 1 : 0;
    }
    // Adam bias corrections of the two drift calls: four double-precision pow() evaluations, one thread each (they only
    // depend on adam_t, which the previous iteration left in shared memory)
    if (tid >= 32 && tid < 36) {
      ChainSm& c = sm.c;
      const int k = tid - 32, t = c.adam_t + 1 + (k >> 1);
      if ((k & 1) == 0) (k == 0 ? c.ss1 : c.ss2) = (float)((double)S->lr / (1.0 - pow((double)kB1, (double)t)));
      else (k == 1 ? c.sq1 : c.sq2) = (float)sqrt(1.0 - pow((double)kB2, (double)t));
    }
    __syncthreads();
    const bool lang = sm.c.is_lang != 0;
    const bool alias = sm.c.alias != 0;
    double a1s[3] = {0.0, 0.0, 0.0}, a2s[3] = {0.0, 0.0, 0.0};
    // ---- first drift call at w (MAIN:473), Adam + noise -> proposal (MAIN:474 / 492)
    if (lang) N::pass(cl, sm, S->Xtrain, ldx, n_train, tr0, tr1, 0, true, buf_d, buf_g);
    N::adam(sm, 1, p_used, k0, k1, step_w, a1s);
    __syncthreads();
    // ---- second drift call at the proposal (MAIN:475): its forward is also the likelihood forward of the proposal (MAIN:505)
    N::pass(cl, sm, S->Xtrain, ldx, n_train, tr0, tr1, 1, lang, buf_d, buf_g);
    const double sse_tr = sm.dsum[0];
    __syncthreads();
    if (lang) N::adam(sm, 2, p_used, k0, k1, step_w, a2s);
    __syncthreads();
    // ---- test forward at the proposal (MAIN:508)
    N::pass(cl, sm, S->Xtest, ldx, n_test, te0, te1, 2, false, buf_d, buf_g);
    const double sse_te = sm.dsum[0];
    __syncthreads();
    // ---- epilogue: BatchNorm-buffer bookkeeping, likelihood, prior, reverse-proposal density, MH test   (run_epilogue)
    {
      const float ub_tr = (float)n_train / (float)(n_train - 1), ub_te = (float)n_test / (float)(n_test - 1);
      double b[3] = {0.0, 0.0, 0.0};   // b_wc, b_wp, b_prior
      if (tid < D3) {
        const int c = tid;
        float nzm[4], nzv[4];
        noise_quad(k0, k1, sm.c.gid, sm.c.step, SEG_BN_RM, (unsigned)(c >> 2), nzm);
        noise_quad(k0, k1, sm.c.gid, sm.c.step, SEG_BN_RV, (unsigned)(c >> 2), nzv);
        const float n_rm = nzm[c & 3] * step_w, n_rv = nzv[c & 3] * step_w;
        const float base_rm = alias ? sm.th[off_rm + c] : sm.wb[c];
        const float base_rv = alias ? sm.th[off_rv + c] : sm.wb[D3 + c];
        float p_rm, p_rv;
        if (lang) {
          const float g_rm = (1.0f - kBnMom) * base_rm + kBnMom * sm.mu[0][c];
          const float g_rv = (1.0f - kBnMom) * base_rv + kBnMom * sm.var[0][c] * ub_tr;
          p_rm = g_rm + n_rm;
          p_rv = g_rv + n_rv;
          if (!alias) {
            const float q_rm = (1.0f - kBnMom) * p_rm + kBnMom * sm.mu[1][c];
            const float q_rv = (1.0f - kBnMom) * p_rv + kBnMom * sm.var[1][c] * ub_tr;
            const double d1 = (double)sm.wb[c] - (double)q_rm, d2 = (double)sm.wb[D3 + c] - (double)q_rv;
            b[0] += d1 * d1 + d2 * d2;
          }
        } else {
          p_rm = base_rm + n_rm;
          p_rv = base_rv + n_rv;
        }
        b[1] += (double)n_rm * n_rm + (double)n_rv * n_rv;
        b[2] += (double)p_rm * p_rm + (double)p_rv * p_rv;
        sm.th[off_rm + c] = (1.0f - kBnMom) * p_rm + kBnMom * sm.mu[2][c];
        sm.th[off_rv + c] = (1.0f - kBnMom) * p_rv + kBnMom * sm.var[2][c] * ub_te;
        sm.wb_tmp[c] = p_rm;
        sm.wb_tmp[D3 + c] = p_rv;
      }
      block_sum<3>(b, sm.red);
      if (tid == 0) {
        ChainSm& c = sm.c;
        float nzn[4];
        noise_quad(k0, k1, c.gid, c.step, SEG_BN_NBT, 0u, nzn);
        const float n_nbt = nzn[0] * step_w;
        const float base_nbt = alias ? sm.th[off_nbt] : sm.wb[2 * D3];
        float p_nbt;
        double nbt_wc = 0.0, nbt_wp = 0.0;
        if (lang) {
          const float g_nbt = truncf(base_nbt) + 1.0f;
          p_nbt = g_nbt + n_nbt;
          nbt_wp = (double)p_nbt - (double)g_nbt;
          if (!alias) {
            const float q_nbt = truncf(p_nbt) + 1.0f;
            nbt_wc = (double)sm.wb[2 * D3] - (double)q_nbt;
          }
        } else {
          p_nbt = base_nbt + n_nbt;
          nbt_wp = (double)p_nbt - (double)base_nbt;
        }
        sm.th[off_nbt] = truncf(p_nbt) + 1.0f;
        sm.wb_tmp[2 * D3] = p_nbt;
        const double sigma_sq = (double)step_w * (double)step_w;
        double diff_prop = 0.0;
        if (lang) {
          const double first = -0.5 * (a2s[2] + b[0] + nbt_wc * nbt_wc) / sigma_sq;        // MAIN:481
          const double second = -0.5 * (a1s[0] + b[1] + nbt_wp * nbt_wp) / sigma_sq;      // MAIN:482
          diff_prop = first - second;
        }
        const double eta_pro = c.eta + c.eta_noise;
        const double tau_pro = exp(eta_pro);
        const double n_tr = (double)n_train * D0, n_te = (double)n_test * D0;
        const double lik_prop = loglik_d(sse_tr, n_tr, tau_pro, c.adapt);
        const double mse_tr = sse_tr / n_tr, mse_te = sse_te / n_te;
        const double sumsq = a1s[1] + b[2] + (double)p_nbt * (double)p_nbt;
        const double prior_prop = -((double)S->w_size * 0.5) * log((double)S->sigma_sq) - sumsq / (2.0 * (double)S->sigma_sq) -
                                  (1.0 + (double)S->nu1) * log(tau_pro) - ((double)S->nu2 / tau_pro);
        const double lik_cur = c.lik;
        const double sum_value = (lik_prop - lik_cur) + (prior_prop - c.prior_cur) + diff_prop;
        const double log_u = log(c.u);
        const bool accept = log_u < sum_value;
        if (accept) {
          c.nacc += 1; c.lik = lik_prop; c.prior_cur = prior_prop; c.eta = eta_pro; c.mse_tr = mse_tr; c.mse_te = mse_te; c.alias = 0;
        } else {
          c.alias = 1;
        }
        if (lang) { c.nlang += 1; c.adam_t += 2; }
        c.last_sse = sse_tr;
        c.prop_sse = sse_tr;
        if (rank == 0 && c.step < S->samples) {
          bae_trace& tr = S->traces[(size_t)chain * S->samples + c.step];
          tr.likelihood_proposal = lik_prop; tr.likelihood = lik_cur; tr.sum_value = sum_value; tr.log_u = log_u;
          tr.prior_proposal = prior_prop; tr.diff_prop = diff_prop; tr.mse_train = c.mse_tr; tr.mse_test = c.mse_te;
          tr.mse_train_proposal = mse_tr; tr.mse_test_proposal = mse_te; tr.eta = c.eta;
          tr.accepted = accept ? 1 : 0; tr.langevin = lang ? 1 : 0; tr.step = c.step;
          for (int k = 0; k < 13; ++k) tr.weights[k] = S->trace_idx_pad[k] >= 0 ? sm.th[S->trace_idx_pad[k]] : 0.f;
        }
        c.step += 1;
        c.accept = accept ? 1 : 0;
      }
      __syncthreads();
      if (sm.c.accept && tid < NB) sm.wb[tid] = sm.wb_tmp[tid];   // accept: `w` buffers <- proposal buffers
      __syncthreads();
    }
  }
  // ---- state back to HBM (cluster rank 0); the peers' shared memory stays mapped until everyone is done
  if (rank == 0) {
    for (int i = tid; i < p_used; i += kT) {
      S->theta[(size_t)chain * S->Pp + i] = sm.th[i];
      S->m[(size_t)chain * S->Pp + i] = sm.m[i];
      S->v[(size_t)chain * S->Pp + i] = sm.v[i];
    }
    if (tid < NB) S->wb[(size_t)chain * S->nbuf + tid] = sm.wb[tid];
    if (tid == 0) {
      const ChainSm& c = sm.c;
      ch.eta[chain] = c.eta; ch.lik[chain] = c.lik; ch.prior_cur[chain] = c.prior_cur; ch.last_sse[chain] = c.last_sse;
      ch.prop_sse[chain] = c.prop_sse; ch.mse_tr[chain] = c.mse_tr; ch.mse_te[chain] = c.mse_te; ch.adapt[chain] = c.adapt;
      ch.adam_t[chain] = c.adam_t; ch.alias[chain] = c.alias; ch.init_count[chain] = c.init_count; ch.nacc[chain] = c.nacc;
      ch.nlang[chain] = c.nlang; ch.step[chain] = c.step; ch.is_lang[chain] = c.is_lang;
    }
  }
  cl.sync();
}

}  // namespace narrow

// ---- host side -------------------------------------------------------------------------------------------------------
// Shapes with a dedicated instantiation (the reference's Swiss Roll network, MAIN:97-100).
bool narrow_supported(const int dims[4], int n_train, int n_test, int p_used) {
  return dims[0] == 3 && dims[1] == 10 && dims[2] == 5 && dims[3] == 2 && p_used <= narrow::kMaxP && n_train <= narrow::kMaxRows &&
         n_test <= narrow::kMaxRows;
}

size_t narrow_smem_bytes() { return sizeof(narrow::Smem<3, 10, 5, 2>); }

// `cluster` CTAs per chain (1, 2, 4 or 8), `n_chains` chains, `n_iter` iterations of MAIN:432-592 each.
cudaError_t launch_narrow(const DevState* dS, int n_chains, int cluster, int n_iter, int p_used, cudaStream_t st) {
  auto fn = narrow::bae_narrow_kernel<3, 10, 5, 2>;
  const size_t smem = narrow_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);   // clusters of 16 CTAs
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_chains * cluster); cfg.blockDim = dim3(narrow::kT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, dS, n_iter, p_used);
}

}  // namespace bae
