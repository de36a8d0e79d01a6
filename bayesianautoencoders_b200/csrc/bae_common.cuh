// Shared host/device definitions for the PT-Langevin sampler (sm_100a).
// Layout, phase program and per-chain state. See DESIGN.md for the data layout in HBM.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/bae_b200.h"

namespace bae {

constexpr int kNumSegs = 17;
// Sorted state_dict key order of the reference's getparameters (MAIN:228-241).
enum SegId : int {
  SEG_BN_BIAS = 0,  // decode.0.bias (beta)
  SEG_BN_NBT,       // decode.0.num_batches_tracked
  SEG_BN_RM,        // decode.0.running_mean
  SEG_BN_RV,        // decode.0.running_var
  SEG_BN_W,         // decode.0.weight (gamma)
  SEG_D1_B, SEG_D1_W,  // decode.1 (layer 4)
  SEG_D3_B, SEG_D3_W,  // decode.3 (layer 5)
  SEG_D5_B, SEG_D5_W,  // decode.5 (layer 6)
  SEG_E0_B, SEG_E0_W,  // encode.0 (layer 1)
  SEG_E2_B, SEG_E2_W,  // encode.2 (layer 2)
  SEG_E4_B, SEG_E4_W   // encode.4 (layer 3)
};

struct Seg {
  int ref_off;    // offset in the reference flat vector
  int pad_off;    // offset in the padded device vector (multiple of 32 floats)
  int rows, cols; // logical shape (1-D tensors: rows = 1)
  int ld;         // padded row pitch in floats (multiple of 4)
  int trainable;  // 0 for the three BatchNorm buffers
};

// ---- GEMM phase descriptors ---------------------------------------------------------------------
enum EpiKind : int {
  EPI_SIGMOID = 0,  // C = sigmoid(acc + bias[c])
  EPI_BIAS,         // C = acc + bias[c]                       (pre-BatchNorm z)
  EPI_LOSS,         // o = acc + bias[c]; d = o - X; partial SSE; C = d * scale (dLoss/dOut)
  EPI_DSIG,         // C = acc * H * (1 - H)                    (delta through a sigmoid)
  EPI_PLAIN,        // C = acc
  EPI_GRADW         // C = acc (weight gradient) + column sums of the A operand -> bias gradient
};

struct GemmDesc {
  const float* A; long long sA; int lda; int aMode;  // aMode 0: A(i,k) = A[i*lda+k]; 1: A(i,k) = A[k*lda+i]
  const float* B; long long sB; int ldb; int bMode;  // bMode 0: B(j,k) = B[j*ldb+k]; 1: B(j,k) = B[k*ldb+j]
  float* C; long long sC; int ldc;
  int M, N, K;
  int epi;
  const float* bias; long long sBias;
  const float* aux; long long sAux; int ldaux;
  float* gbias; long long sGb;
  double* part; int partStride;  // [chain][partStride][2] loss partials (sum d^2, sum d)
  float scale;
  int needLang;  // 1: only chains whose current step is a Langevin step
  int aAct, bAct, cAct, auxAct;   // operand is an activation / delta buffer: indexed by the chain's slot in the resident GROUP
                                  // (chain - c_begin of the launch), not by the chain (DevState::G)
  int tm, tn;    // tiles along M and N
  int big;       // 1: 128x128 tiles, 0: 64x64 tiles
  int ksplit;    // weight-gradient GEMMs: the contraction K (= data rows) is cut into this many segments.
                 // FP32 path: one work item per segment, segment s writes gradient copy s (stride = C_alloc * Pp floats),
                 //            summed in fixed order by the Adam phase.
                 // tcgen05 path: the segments of a tile run back to back on one CTA pair and accumulate into ONE copy.
  // ---- tcgen05 path (3xTF32 with the hi/lo split done in shared memory)
  int tcBN;             // tile width (multiple of 16; multiple of 32 when B is N-major), <= 256
  int tcBBytes;         // bytes of one B tile (hi or lo) per K block
  int tcMain;           // TMEM accumulators the hi*hi term rotates over; (tcMain + 1) * tcBN <= 512 columns
  int tmIdx[3];         // tensor maps: A, B (loads), C (32 x 32 store boxes of the epilogue)
  unsigned idesc;       // tcgen05 instruction descriptor
  int biasCol;          // EPI_GRADW: output column that holds the bias gradient (ones column trick), or -1
  // ---- 3xFP16 split (DevState::f16): where the power-of-two scale of each operand comes from (bae_tc.cuh)
  int aSc, bSc;         // >= 0: amax slot (AmaxSlot, per chain); SC_UNIT: values in [0, 1]; SC_XTRAIN / SC_XTEST: DevState::xexp
  int cSc;              // amax slot the epilogue folds max |C| into (deltas), or -1
  // ---- pre-split operand images (3xFP16 split, DevState::preconv; layout: bae_tc.cuh "operand images"). Null: the converter
  // warps split this operand's FP32 tile in shared memory. Non-null (weights: once per proposal; data: once per upload): the
  // producer copies the FP16 hi / lo tiles of a K block straight into the converted stage.
  const unsigned char* aImg; long long sAImg; int aKbs;   // base, bytes per chain, bytes per K block (hi plane; lo plane at aKbs / 2)
  const unsigned char* bImg; long long sBImg; int bKbs;
  // Forward GEMMs over the training matrix whose A operand is an activation (H1, H2, Y, H4, H5): the converter group that split
  // a K block of the A tile also sends the converted tiles to the MN-major image of that activation (tensor maps tmOut: hi plane,
  // tmOut + 1: lo plane), which the weight-gradient GEMM of the same layer reads as its B operand (indexed by the chain's group
  // slot). Only tile_n == 0 items store. The A tensor map of such a GEMM reaches one column further: the ones column.
  unsigned char* aOut; int tmOut;
  int aOnes;            // 1: the A tensor map reaches one column further than K (the activation's ones column goes into the image)
};

// Pre-split images of one weight matrix inside a chain's image slab (DevState::wimg): K-major (forward GEMMs: contraction over
// the matrix's columns) and MN-major (delta GEMMs: contraction over its rows; offM < 0: not needed)
struct WImg {
  int seg, items;          // tensor; work items of bae_wconv_kernel (8 rows x 64 units of 8 columns)
  long long offK, offM;    // byte offsets inside the slab
  int kbsK, kbsM;          // bytes per K block
};

// Magnitude bounds behind the FP16 operand scales: one float per (slot, chain), DevState::amax[slot * C_alloc + chain], folded
// with atomicMax on the bit pattern (non-negative floats order like unsigned integers) and zeroed by bae_amax_zero_kernel
// before the phase range that produces them.
enum AmaxSlot : int {
  AMAX_W = 0,                          // the six weight matrices of the chain's live model (PH_WAMAX)
  AMAX_Y,                              // BatchNorm output of the most recent forward pass (bound derived by run_bn_fwd_wide)
  AMAX_DOUT, AMAX_D5, AMAX_D4, AMAX_DY, AMAX_D2, AMAX_D1,   // deltas (GEMM epilogues; DY: bound derived by the BatchNorm backward)
  AMAX_DYIN,                           // dy as the GEMM epilogue wrote it, in front of the BatchNorm backward
  kAmaxSlots
};
enum ScaleSrc : int { SC_UNIT = -1, SC_XTRAIN = -2, SC_XTEST = -3 };
constexpr int kF16TopExp = 13;   // operands are scaled so that their bound lands in [2^13, 2^14)

// exponent e with bound * 2^e in [2^13, 2^14), clamped so that sums of two exponents (+ 11) stay inside the FP32 range
__host__ __device__ inline int f16_scale_exp(float bound) {
  unsigned u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(bound);
#else
  memcpy(&u, &bound, 4);
#endif
  const int e = kF16TopExp - ((int)((u >> 23) & 0xFFu) - 127);
  return e < -50 ? -50 : (e > 50 ? 50 : e);
}

enum PhaseKind : int {
  PH_PROLOGUE = 0, PH_GEMM, PH_BN_FWD, PH_BN_BWD, PH_ADAM1, PH_ADAM2, PH_EPILOGUE,
  PH_SWAP_DECIDE, PH_SWAP_GATHER, PH_SWAP_SCATTER,
  PH_WAMAX        // 3xFP16 split: max |w| over the weight matrices of every chain (after whatever rewrote theta)
};

struct PhaseDesc {
  int kind;
  int g0, g1;   // GEMM descriptor indices (g1 = -1 if single)
  int arg;      // BN: stats slot (0 first train pass, 1 second train pass, 2 test); rows via argRows; PH_WAMAX: 1 = only after an exchange round
  int argRows;  // BN: number of data rows
  int needLang;
  int sched_off, sched_max;   // tcgen05 path: this phase's item schedule inside DevState::sched ([pair][sched_max], -1 padded), or -1
};

constexpr int kMaxGemm = 40;
constexpr int kMaxPhase = 64;
constexpr int kAdamChunk = 8192;   // floats per Adam work item (256 threads x 8 float4)
constexpr int kThreads = 256;
constexpr int kAuxAdamCtas = 4;      // bae_aux_kernel<AUX_ADAM>: 256-thread CTAs per SM
constexpr int kAuxBnCtas = 2;        // bae_aux_kernel<AUX_BN>

// Per-chain scalar state (structure of arrays, length C_alloc = C + 1; slot C is the eval scratch chain).
struct ChainArrays {
  double *eta, *lik, *prior_cur, *last_sse, *prop_sse, *mse_tr, *mse_te, *T, *adapt;
  int *adam_t, *alias, *init_count, *nacc, *nlang, *step;
  int* gid;   // global chain id (keys the Philox streams): chain_id_base + chain, or the ladder position of a sharded run
  // per-step scratch
  int* is_lang;
  double *lx, *eta_noise, *u;
  float *ss1, *sq1, *ss2, *sq2;  // Adam step_size and sqrt(bias_correction2) for the two drift calls
  int* plain;                    // 1: this iteration's drift is w - drift_a * grad - drift_b * w (SGD after pt / MALA), no Adam state
  float *drift_a, *drift_b;
};

struct DevState {
  // ---- static configuration
  int d0, d1, d2, d3;
  int ld0, ld1, ld2, ld3;
  int n_train, n_test, n_max;
  int C, C_alloc, n_rungs, chain_id_base;
  int G;              // chains whose activation buffers are resident at a time (chain groups; G == C: all of them)
  int samples, pt_step, swap_interval, use_langevin, drift_mode;
  float lr, step_w, step_eta, l_prob, sigma_sq, nu1, nu2;
  unsigned seed_lo, seed_hi;
  int w_size, Pp, n_adam_items, nbuf;  // nbuf = 2*d3 + 1
  Seg seg[kNumSegs];
  int seg_ksplit[kNumSegs];   // gradient copies to sum per tensor (1 = written directly)
  int max_ksplit;
  // ---- chain-batched vectors [C_alloc][Pp]
  float *theta, *m, *v, *wcur, *grad;
  // ---- BatchNorm side state
  float* wb;        // [C_alloc][nbuf]   buffers of the detached `w` dict: rm[d3], rv[d3], nbt
  float* wb_tmp;    // swap scratch
  float* bn_mu;     // [3][C_alloc][ld3] batch mean per stats slot
  float* bn_var;    // [3][C_alloc][ld3] biased batch variance
  float* bn_inv;    // [C_alloc][ld3]    1/sqrt(var+eps) of the most recent forward
  // ---- data + activations
  float *Xtrain, *Xtest;
  float *H1, *H2, *Z, *Y, *H4, *H5, *DOUT, *D5, *D4, *DY, *D2, *D1;
  // ---- partial sums
  double* loss_part;  // [3][C_alloc][loss_stride][2]
  int loss_stride;
  double* adam_part;  // [C_alloc][n_adam_items][3]  (sum noise^2, sum prop^2, sum wc_delta^2)
  // ---- per-chain scalars
  ChainArrays ch;
  // ---- traces
  bae_trace* traces;  // [C][samples]
  int trace_idx_pad[13];  // padded index of the 13 traced entries (-1: beyond w_size)
  // ---- swap
  int* swap_src;     // [C] chain whose configuration ends at position p
  int* swap_flag;    // [C] 1 if position p receives a different configuration
  double* swap_tmp;  // [C][2] eta, last_sse scratch
  long long *num_swap, *total_swap;
  int swap_round;
  // ---- ladder sharded over ranks (bae_shard_config / bae_swap_exchange); world == 1: not sharded
  int shard_rank, shard_world;
  double* xt_local;   // [C][5]         (eta, stored likelihood, adapttemp, SSE of w, alias flag) of the resident chains
  double* xt_all;     // [world][C][5]  all-gathered
  int* xsrc;          // [E][R]         global rung whose configuration ends at rung p (identical on every rank)
  // ---- tcgen05 path
  int tc;             // 1: GEMM phases run on the tensor cores; row pitches reserve a ones column
  int f16;            // 1: 3xFP16 split (kind::f16, K = 16 per MMA, scaled operands); 0: 3xTF32 split (kind::tf32)
  float* amax;        // [kAmaxSlots][C_alloc] magnitude bounds of the scaled operands (AmaxSlot)
  int xexp[2];        // scale exponents of the training / test matrix (host scan at bae_upload_data)
  int preconv;        // 1: weights and data are read as pre-split FP16 images (GemmDesc::aImg / bImg), rewritten by bae_wconv_kernel
  unsigned char* wimg;        // [C_alloc][wimg_stride] bytes
  long long wimg_stride;
  int wimg_items;             // work items of bae_wconv_kernel per chain
  unsigned char* himg[5];     // MN-major images of the activations H1, H2, Y, H4, H5 of the resident chain group: [G][himg_stride[i]]
  long long himg_stride[5];
  int himg_kbs[5];
  WImg wi[6];
  void* tmaps;        // CUtensorMap array in global memory (64-byte aligned)
  int* sched;         // longest-first item schedules of the GEMM phases (PhaseDesc::sched_off), built for `sched_chains` chains
  int sched_chains, sched_pairs;   // ... per launch and `sched_pairs` CTA pairs
  int dbg_flags;      // BAE_TC_DBGFLAGS: timing experiments only (results are wrong when set), see bae_tc.cuh
  int dbg_phase;      // BAE_TC_TRACE_PHASE: trace only this phase index (-1: every GEMM phase, the last one stays)
  long long* dbg;     // optional timeline trace of CTA 0 (BAE_TC_TRACE=1): [role][event] clock64 stamps
  // ---- program
  int n_gemm, n_phase;
  GemmDesc gemm[kMaxGemm];
  PhaseDesc phase[kMaxPhase];
};

// ---- Philox4x32-10 ----------------------------------------------------------------------------
// Stream ids (counter word 1 bits 8..15); low 8 bits carry the tensor id for weight noise.
enum PhiloxStream : unsigned { STREAM_WEIGHT_NOISE = 1, STREAM_STEP_SCALARS = 2, STREAM_SWAP = 3 };

struct u4 { unsigned x, y, z, w; };

__host__ __device__ inline void mulhilo(unsigned a, unsigned b, unsigned& hi, unsigned& lo) {
  unsigned long long p = (unsigned long long)a * (unsigned long long)b;
  hi = (unsigned)(p >> 32);
  lo = (unsigned)p;
}

__host__ __device__ inline u4 philox4x32_10(u4 c, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c.x, hi0, lo0);
    mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
    u4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// uniform in (0,1): ((x >> 8) + 0.5) * 2^-24
__host__ __device__ inline float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

}  // namespace bae
