// C-ABI of the sampler (include/bae_b200.h): host-side state, phase-program construction, launches.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_profiler_api.h>
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <map>

#include "bae_common.cuh"

namespace bae {
int max_coop_blocks_per_sm(int use_tc);
int max_coop_grid_tc();
cudaError_t launch_program(const DevState* dS, int use_tc, int grid, int ph_begin, int ph_end, int n_rep, int c_begin,
                           int c_end, int force_swap, cudaStream_t st);
cudaError_t launch_phase(const DevState* dS, int use_tc, int grid, int pi, int c_begin, int c_end, int force_swap,
                         int epi_a, int epi_b, int front_mode, cudaStream_t st);
int aux_blocks_per_sm(int kind);
cudaError_t launch_aux(const DevState* dS, int kind, int grid, int pi, int c_begin, int c_end, int force_swap,
                       cudaStream_t st);
cudaError_t launch_split(const float* src, float* hi, float* lo, size_t n, cudaStream_t st);
cudaError_t launch_grad_reduce(const DevState* dS, int chain, cudaStream_t st);
cudaError_t launch_debug_draws(const DevState* dS, int chain, int step, float* noise, double* scalars, cudaStream_t st);
cudaError_t launch_debug_swap_u(const DevState* dS, int ens, int round, int pair, double* out, cudaStream_t st);
cudaError_t launch_scalars(const DevState* dS, int chain, bae_chain_scalars* staged, int put, cudaStream_t st);
bool narrow_supported(const int dims[4], int n_train, int n_test, int p_used);
cudaError_t launch_narrow(const DevState* dS, int n_chains, int cluster, int n_iter, int p_used, cudaStream_t st);
cudaError_t launch_swap_cta(const DevState* dS, int force_swap, cudaStream_t st);
cudaError_t launch_amax_zero(const DevState* dS, unsigned mask, int c_begin, int c_end, int gated, cudaStream_t st);
cudaError_t launch_wconv(const DevState* dS, int grid, int c_begin, int c_end, int gated, cudaStream_t st);
cudaError_t launch_himg_ones(unsigned char* img, long long stride, int kbs, int d, int rows, int nslots, cudaStream_t st);
cudaError_t launch_ximg(const float* src, int ld, int R, int ccK, int ccM, int e, unsigned char* imgK, int kbsK, unsigned char* imgM,
                        int kbsM, cudaStream_t st);
cudaError_t launch_xtable(const DevState* dS, cudaStream_t st);
cudaError_t launch_xsweep(const DevState* dS, cudaStream_t st);
cudaError_t launch_xgather(const DevState* dS, int grid, cudaStream_t st);
cudaError_t launch_xscatter(const DevState* dS, int grid, cudaStream_t st);
cudaError_t launch_extract(const float* src, int ld, const float* X, int ldx, int rows, int d, float inv_scale, int mode,
                           float* out, cudaStream_t st);
cudaError_t launch_pack(const DevState* dS, const float* ref_vec, float* pad_vec, cudaStream_t st);
cudaError_t launch_unpack(const DevState* dS, const float* pad_vec, float* ref_vec, cudaStream_t st);
}  // namespace bae

using namespace bae;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess)                                                                          \
      return fail(BAE_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));                   \
  } while (0)

// Phase index map of the step program (see build_program).
enum PhaseIdx : int {
  P_PRO = 0,
  P_WAMAX0 = 1,  // 3xFP16 split: weight bounds after an exchange round (gated) / forced by the host in front of a program
  P_FWD1 = 2,    // 7 phases
  P_BWD1 = 9,    // 7 phases
  P_ADAM1 = 16,  // (3xFP16 split: also folds the weight bounds of the proposal it writes)
  P_FWD2 = 17,
  P_BWD2 = 24,
  P_ADAM2 = 31,
  P_FWDT = 32,
  P_EPI = 39,
  P_SWAP = 40,   // decide, gather, scatter
  P_END = 43
};

struct bae_handle {
  bae_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  DevState hs;            // host copy (device pointers inside)
  DevState* dS = nullptr;
  std::vector<void*> allocs;
  int grid = 0;
  int grid_hint = 296;    // CTAs of the persistent kernel (used to size K splits before the grid is known)
  bool data_loaded = false;
  std::vector<CUtensorMap> tmaps_host;
  std::vector<float> ones_col;
  std::string tmap_error;
  float* stage_ref = nullptr;    // [w_size] device staging, reference order
  float* stage_noise = nullptr;  // [w_size]
  double* stage_d = nullptr;     // small device scratch for doubles
  bae_chain_scalars* stage_sc = nullptr;   // device staging of one chain's scalars (bae_get_state / bae_set_state)
  float* stage_rows = nullptr;             // [n_max][max(d0, d3)] dense staging of bae_forward_all (allocated on first use)
  int32_t* xsrc_host = nullptr;            // pinned [E][R]: the permutation of the last sharded exchange round
  std::vector<int32_t> xops;
  int64_t vectors_sent = 0;
  // pre-split images of the data matrices (DevState::preconv): K-major of train / test (first forward GEMM), MN-major of train
  // (weight gradient of encode.0, ones column included)
  unsigned char* ximgK[2] = {nullptr, nullptr};
  unsigned char* ximgM = nullptr;
  int ximgK_kbs[2] = {0, 0}, ximgM_kbs = 0;
  int n_sms = 148;
  // narrow-width step kernel (bae_narrow.cu): one thread-block cluster per chain
  bool narrow = false;
  int narrow_cluster = 1, p_used = 0;
  long long host_step = 0;   // iterations executed since the chains were initialised (all chains advance together)
  int64_t launches = 0;
  int64_t n_emit = 0;          // kernels issued by launch_one (a phase may come with an amax-zeroing launch, or with none)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  bool phase_launch_env = false;
  bool phase_launch = false;   // one launch per phase (default on the tcgen05 path) instead of the persistent kernel
  bool use_graph = true;       // phase launches of a call are replayed from a CUDA graph
  int grid_aux[16] = {};       // CTAs of bae_aux_kernel per phase kind (non-GEMM phases of the tcgen05 path)
  std::map<std::array<int, 4>, cudaGraphExec_t> graphs;   // (p_begin, p_end, n_rep, force_swap) -> instantiated graph
  std::map<std::array<int, 4>, int> graph_launches;       // kernel nodes of each graph
  // BAE_PHASE_TIMES=1 (debug): direct launches with an event after every phase; bae_debug_phase_times sums the intervals
  bool phase_times = false;
  std::vector<std::pair<int, cudaEvent_t>> pt_events;   // (phase index or -1 = start marker, event)
};

template <typename T>
static cudaError_t dalloc(bae_handle* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(q, 0, n * sizeof(T) + 256, h->stream);
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return e;
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static void build_layout(DevState& s) {
  const int d0 = s.d0, d1 = s.d1, d2 = s.d2, d3 = s.d3;
  struct Sh { int rows, cols, tr; };
  const Sh sh[kNumSegs] = {
      {1, d3, 1}, {1, 1, 0}, {1, d3, 0}, {1, d3, 0}, {1, d3, 1},  // BN bias, nbt, rm, rv, weight
      {1, d2, 1}, {d2, d3, 1},                                      // decode.1
      {1, d1, 1}, {d1, d2, 1},                                      // decode.3
      {1, d0, 1}, {d0, d1, 1},                                      // decode.5
      {1, d1, 1}, {d1, d0, 1},                                      // encode.0
      {1, d2, 1}, {d2, d1, 1},                                      // encode.2
      {1, d3, 1}, {d3, d2, 1}};                                     // encode.4
  int ref = 0, pad = 0;
  for (int i = 0; i < kNumSegs; ++i) {
    Seg& g = s.seg[i];
    g.rows = sh[i].rows;
    g.cols = sh[i].cols;
    g.ld = round_up(sh[i].cols, 4);
    g.trainable = sh[i].tr;
    g.ref_off = ref;
    g.pad_off = pad;
    ref += g.rows * g.cols;
    pad += round_up(g.rows * g.ld, 32);
  }
  s.w_size = ref;
  s.Pp = round_up(pad, kAdamChunk);
  s.n_adam_items = s.Pp / kAdamChunk;
  s.nbuf = 2 * d3 + 1;
  const int tidx[13] = {0, 100, 10000, 5000, 2000, 3000, 4000, 5000, 6000, 7000, 8000, 9000, 11000};  // MAIN:568-592
  for (int k = 0; k < 13; ++k) {
    s.trace_idx_pad[k] = -1;
    if (tidx[k] >= s.w_size) continue;
    for (int i = 0; i < kNumSegs; ++i) {
      const Seg& g = s.seg[i];
      if (tidx[k] >= g.ref_off && tidx[k] < g.ref_off + g.rows * g.cols) {
        int e = tidx[k] - g.ref_off;
        s.trace_idx_pad[k] = g.pad_off + (e / g.cols) * g.ld + (e % g.cols);
      }
    }
  }
}

static int tc_max_bn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BAE_TC_MAXBN");
    v = e ? std::max(32, std::min(256, atoi(e) / 16 * 16)) : 256;
  }
  return v;
}

// Pre-split operand image (bae_tc.cuh "operand images") of a matrix used with `extent` tile rows / MN and a contraction of length K:
// bytes per K block (hi plane + lo plane, the extent padded to whole 256-wide tiles) and in total
static int img_kbs(int extent) { return 64 * round_up(extent, 256); }
static size_t img_bytes(int extent, int K) { return (size_t)((K + 15) / 16) * (size_t)img_kbs(extent); }

static GemmDesc make_gemm(const DevState& s, int M, int N, int K, int b_nmajor = 0) {
  GemmDesc g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K;
  g.biasCol = -1;
  g.ksplit = 1;
  if (s.tc) {
    // tcgen05 CTA-pair tiles: 256 x BN, BN <= 256 (a multiple of 16) chosen so the N tiles are balanced. Each CTA of a
    // pair holds BN/2 columns of B; an N-major B operand is loaded in 32-column chunks, so the half tile in shared
    // memory may be wider than BN/2.
    const int maxbn = tc_max_bn();
    g.tm = (M + 255) / 256;
    g.tn = (N + maxbn - 1) / maxbn;
    g.tcBN = round_up((N + g.tn - 1) / g.tn, 16);
    const int half = g.tcBN / 2;
    g.tcBBytes = b_nmajor ? round_up(half, 32) / 32 * 2048 : half * 64;   // chunks of [16 k][32 n] : rows of 64 B
    g.tcMain = 1;
    return g;
  }
  g.big = (M >= 128 && N > 64) ? 1 : 0;
  const int bt = g.big ? 128 : 64;
  g.tm = (M + bt - 1) / bt;
  g.tn = (N + bt - 1) / bt;
  return g;
}

// FP32 path, weight-gradient GEMMs: few chains / few output tiles but a long contraction (K = rows). Split K over
// work items so the phase fills the grid; the copies are summed in fixed order (deterministic) by the Adam phase.
static int grad_ksplit(const DevState& s, int grid_hint, const GemmDesc& g, int rows) {
  if (s.tc) {
    // tcgen05 path: the tensor core truncates when it adds into its accumulator, so the accumulation chains stay short:
    // 3xTF32 (K = 8 per MMA): the length of a forward layer's (<= ~640 rows, 80 accumulations); 3xFP16 (K = 16 per MMA):
    // 1000 rows = 63 accumulations. Every segment costs a drain and an epilogue pass (a reduce-add of the whole tile): at
    // Madelon 2 segments of 1000 rows instead of 4 of 500 are worth 4.4% of the step, for a gradient error of 0.78e-5
    // instead of 0.73e-5 at iteration 0 (tolerance 1.2e-5; profiles/r2_parity_measured.md)
    const char* e = getenv("BAE_TC_KSPLIT_ROWS");
    const int per = e ? std::max(64, atoi(e)) : (s.f16 ? 1000 : 640);
    return std::max(1, (rows + per - 1) / per);   // segments run back to back on one CTA pair and accumulate into one copy
  }
  const int tiles = g.tm * g.tn * s.C;
  const int nk = (rows + 15) / 16;
  const int ks = (2 * grid_hint + tiles - 1) / tiles;
  return std::max(1, std::min(std::min(ks, 16), nk / 8));
}

// ---- TMA tensor maps (driver entry point fetched at run time: no link dependency on libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// K-major operand: element (row, k) at ptr[chain*cs + row*ld + k]; box = 16 k x box_rows rows
static int tmap_kmajor(CUtensorMap* out, const float* ptr, int K, int rows, int ld, int nchain, long long cs, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)std::max(1, nchain)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(cs > 0 ? cs : (long long)rows * ld) * 4};
  cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};   // 16 floats of K per pipeline stage (64-byte rows)
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// MN-major operand: element (t, k) at ptr[chain*cs + k*ld + t]; one box = [16 k][32 t] (a 2 KB swizzled chunk);
// the producer issues one TMA per 32-wide chunk of the tile dimension.
static int tmap_mnmajor(CUtensorMap* out, const float* ptr, int extent, int K, int ld, int nchain, long long cs, int /*box_chunks*/) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)extent, (cuuint64_t)K, (cuuint64_t)std::max(1, nchain)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(cs > 0 ? cs : (long long)K * ld) * 4};
  cuuint32_t box[3] = {32, 16, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Output of a GEMM: element (row, col) of slice z at ptr[z*cs + row*ld + col]; the epilogue stores 32 x 32 boxes from a
// 128-byte-swizzled staging tile. Rows >= `rows` and columns >= `cols` of a box are clipped by the copy engine.
static int tmap_store(CUtensorMap* out, const float* ptr, int cols, int rows, int ld, int nslice, long long cs) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)std::max(1, nslice)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(cs > 0 ? cs : (long long)rows * ld) * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

// index of a weight matrix in DevState::wi
static const int kWimgSegs[6] = {SEG_E0_W, SEG_E2_W, SEG_E4_W, SEG_D1_W, SEG_D3_W, SEG_D5_W};
static int wimg_index(int seg) {
  for (int i = 0; i < 6; ++i)
    if (kWimgSegs[i] == seg) return i;
  return 0;
}

// Plane (hi or lo) of an MN-major activation image as the target of the converters' box stores: a converted K-major A tile
// [row group 2 kbrow + jrow][k half jc][r][16 B] is the box (jc and r: 256 contiguous bytes at column unit 2 kb; jrow; 8 kbrow)
static int tmap_image_store(CUtensorMap* out, unsigned char* plane, int kbs, int rows, int nslot, long long slot_stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  const cuuint64_t t8 = (cuuint64_t)kbs / 512;
  cuuint64_t dims[4] = {t8 * 64, 2, (cuuint64_t)((rows + 15) / 16), (cuuint64_t)std::max(1, nslot)};
  cuuint64_t strides[3] = {t8 * 128, (cuuint64_t)kbs, (cuuint64_t)slot_stride};
  cuuint32_t box[4] = {128, 2, 8, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)plane, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

static void build_program(bae_handle* h) {
  DevState& s = h->hs;
  const long long P = s.Pp;
  const long long a1 = (long long)s.n_max * s.ld1, a2 = (long long)s.n_max * s.ld2, a3 = (long long)s.n_max * s.ld3,
                  a0 = (long long)s.n_max * s.ld0;
  const bool tcm = s.tc != 0;
  auto W = [&](int seg) { return s.theta + s.seg[seg].pad_off; };
  auto Wraw = W;
  auto G = [&](int seg) { return s.grad + s.seg[seg].pad_off; };
  h->tmaps_host.clear();
  h->tmap_error.clear();
  // registers the two tensor maps (A, B) of a GEMM
  auto add_maps = [&](GemmDesc& g) {
    if (!tcm) return;
    const int nchA = g.sA ? (g.aAct ? s.G : s.C_alloc) : 1, nchB = g.sB ? (g.bAct ? s.G : s.C_alloc) : 1;
    const float* ptrs[2] = {g.A, g.B};
    for (int i = 0; i < 2; ++i) {
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      int rc;
      if (i < 1) {
        rc = g.aMode == 0 ? tmap_kmajor(&tm, ptrs[i], g.K + (g.aOnes ? 1 : 0), g.M, g.lda, nchA, g.sA, 128)
                          : tmap_mnmajor(&tm, ptrs[i], g.M, g.K, g.lda, nchA, g.sA, 4);
      } else {
        rc = g.bMode == 0 ? tmap_kmajor(&tm, ptrs[i], g.K, g.N, g.ldb, nchB, g.sB, g.tcBN / 2)
                          : tmap_mnmajor(&tm, ptrs[i], g.N, g.K, g.ldb, nchB, g.sB, 0);
      }
      if (rc != 0 && h->tmap_error.empty())
        h->tmap_error = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string(rc) + ") for GEMM operand " + std::to_string(i);
      g.tmIdx[i] = (int)h->tmaps_host.size();
      h->tmaps_host.push_back(tm);
    }
    {
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      const bool gw = g.epi == EPI_GRADW;
      const int cols = (gw && g.biasCol >= 0) ? std::min(g.N, g.biasCol) : g.N;
      const int nslice = g.sC ? (g.cAct ? s.G : (gw ? std::max(1, s.max_ksplit) : 1) * s.C_alloc) : 1;
      const int rc = tmap_store(&tm, g.C, cols, g.M, g.ldc, nslice, g.sC);
      if (rc != 0 && h->tmap_error.empty())
        h->tmap_error = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string(rc) + ") for a GEMM output";
      g.tmIdx[2] = (int)h->tmaps_host.size();
      h->tmaps_host.push_back(tm);
    }
    // instruction descriptor of the pair's MMA (cta_group::2): M = 256, N = tcBN
    // (D format F32 at bit 4; A / B formats at bits 7 / 10: 2 = TF32 for kind::tf32, 0 = F16 for kind::f16)
    const unsigned fmt = s.f16 ? 0u : ((2u << 7) | (2u << 10));
    g.idesc = (1u << 4) | fmt | ((unsigned)g.aMode << 15) | ((unsigned)g.bMode << 16) |
              ((unsigned)(g.tcBN >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
  };
  // registers the two box-store maps (hi plane, lo plane) of an image the converters of `g` persist their A tiles into
  auto add_out_maps = [&](GemmDesc& g, unsigned char* img, int kbs, long long stride) {
    g.aOut = img;
    g.tmOut = (int)h->tmaps_host.size();
    for (int pl = 0; pl < 2; ++pl) {
      CUtensorMap tm;
      memset(&tm, 0, sizeof(tm));
      const int rc = tmap_image_store(&tm, img + (size_t)pl * (kbs / 2), kbs, s.n_train, s.G, stride);
      if (rc != 0 && h->tmap_error.empty())
        h->tmap_error = "cuTensorMapEncodeTiled failed (CUresult " + std::to_string(rc) + ") for an operand image";
      h->tmaps_host.push_back(tm);
    }
  };
  int ng = 0;
  // ---- forward descriptors for (train, first pass), (train, second pass), (test)
  int fwd_base[3];
  for (int pass = 0; pass < 3; ++pass) {
    const float* Xraw = (pass == 2) ? s.Xtest : s.Xtrain;
    const float* X = Xraw;
    const int N = (pass == 2) ? s.n_test : s.n_train;
    const int need = (pass == 0) ? 1 : 0;
    fwd_base[pass] = ng;
    struct L { const float* A; long long sA; int lda; int wseg, bseg; float* C; long long sC; int ldc; int din, dout; int epi; };
    const L ls[6] = {
        {X, 0, s.ld0, SEG_E0_W, SEG_E0_B, s.H1, a1, s.ld1, s.d0, s.d1, EPI_SIGMOID},
        {s.H1, a1, s.ld1, SEG_E2_W, SEG_E2_B, s.H2, a2, s.ld2, s.d1, s.d2, EPI_SIGMOID},
        {s.H2, a2, s.ld2, SEG_E4_W, SEG_E4_B, s.Z, a3, s.ld3, s.d2, s.d3, EPI_BIAS},
        {s.Y, a3, s.ld3, SEG_D1_W, SEG_D1_B, s.H4, a2, s.ld2, s.d3, s.d2, EPI_SIGMOID},
        {s.H4, a2, s.ld2, SEG_D3_W, SEG_D3_B, s.H5, a1, s.ld1, s.d2, s.d1, EPI_SIGMOID},
        {s.H5, a1, s.ld1, SEG_D5_W, SEG_D5_B, s.DOUT, a0, s.ld0, s.d1, s.d0, EPI_LOSS}};
    for (int l = 0; l < 6; ++l) {
      GemmDesc g = make_gemm(s, N, ls[l].dout, ls[l].din, 0);
      g.A = ls[l].A; g.sA = ls[l].sA; g.lda = ls[l].lda; g.aMode = 0;
      g.B = W(ls[l].wseg); g.sB = P; g.ldb = s.seg[ls[l].wseg].ld; g.bMode = 0;
      g.C = ls[l].C; g.sC = ls[l].sC; g.ldc = ls[l].ldc;
      g.aAct = ls[l].sA != 0; g.cAct = 1;      // activations: slot of the resident chain group
      g.epi = ls[l].epi;
      g.bias = Wraw(ls[l].bseg); g.sBias = P;
      g.needLang = need;
      // 3xFP16 split: bounds of the operands (data: host scan; sigmoid outputs: 1; BatchNorm output: tracked; weights: tracked)
      g.aSc = (l == 0) ? (pass == 2 ? (int)SC_XTEST : (int)SC_XTRAIN) : (l == 3 ? (int)AMAX_Y : (int)SC_UNIT);
      g.bSc = AMAX_W;
      g.cSc = (ls[l].epi == EPI_LOSS) ? AMAX_DOUT : -1;
      if (s.preconv) {   // weights and data arrive pre-split
        const WImg& wi = s.wi[wimg_index(ls[l].wseg)];
        g.bImg = s.wimg + wi.offK; g.sBImg = s.wimg_stride; g.bKbs = wi.kbsK;
        if (l == 0) { g.aImg = h->ximgK[pass == 2 ? 1 : 0]; g.sAImg = 0; g.aKbs = h->ximgK_kbs[pass == 2 ? 1 : 0]; }
        // the activation a pass over the training matrix reads as its A operand is persisted as an image by the converter warps
        if (pass != 2 && l >= 1 && s.himg[l - 1]) {
          add_out_maps(g, s.himg[l - 1], s.himg_kbs[l - 1], s.himg_stride[l - 1]);
          g.aOnes = 1;
        }
      }
      if (ls[l].epi == EPI_LOSS) {
        g.aux = Xraw; g.sAux = 0; g.ldaux = s.ld0;
        g.scale = 2.0f / ((float)N * (float)s.d0);   // d MSELoss / d out (mean over all N*d0 elements, MAIN:150,223)
        g.part = s.loss_part + (size_t)pass * s.C_alloc * s.loss_stride * 2;
        g.partStride = s.loss_stride;
      }
      add_maps(g);
      s.gemm[ng++] = g;
    }
  }
  // ---- backward descriptors (shared by both drift calls)
  const int N = s.n_train;
  const float* Xtr = s.Xtrain;
  struct BL { float* dOut; long long sD; int ldd; int dout; const float* Ain; long long sAin; int ldain; int din;
              int wseg, bseg; float* dIn; long long sDin; int lddin; const float* Hin; int epi_x; };
  const BL bl[6] = {
      {s.DOUT, a0, s.ld0, s.d0, s.H5, a1, s.ld1, s.d1, SEG_D5_W, SEG_D5_B, s.D5, a1, s.ld1, s.H5, EPI_DSIG},
      {s.D5, a1, s.ld1, s.d1, s.H4, a2, s.ld2, s.d2, SEG_D3_W, SEG_D3_B, s.D4, a2, s.ld2, s.H4, EPI_DSIG},
      {s.D4, a2, s.ld2, s.d2, s.Y, a3, s.ld3, s.d3, SEG_D1_W, SEG_D1_B, s.DY, a3, s.ld3, nullptr, EPI_PLAIN},
      {s.DY, a3, s.ld3, s.d3, s.H2, a2, s.ld2, s.d2, SEG_E4_W, -1 /* encode.4.bias gradient is exactly 0 */, s.D2, a2, s.ld2, s.H2, EPI_DSIG},
      {s.D2, a2, s.ld2, s.d2, s.H1, a1, s.ld1, s.d1, SEG_E2_W, SEG_E2_B, s.D1, a1, s.ld1, s.H1, EPI_DSIG},
      {s.D1, a1, s.ld1, s.d1, Xtr, 0, s.ld0, s.d0, SEG_E0_W, SEG_E0_B, nullptr, 0, 0, nullptr, EPI_PLAIN}};
  int bw_idx[6], bx_idx[6];
  for (int l = 0; l < 6; ++l) {
    // dW[o,i] = sum_n dOut[n,o] * Ain[n,i]; on the tcgen05 path a ones column appended to Ain yields the bias gradient
    const bool ones = tcm && bl[l].bseg >= 0;
    GemmDesc g = make_gemm(s, bl[l].dout, bl[l].din + (ones ? 1 : 0), N, 1);
    g.A = bl[l].dOut; g.sA = bl[l].sD; g.lda = bl[l].ldd; g.aMode = 1;
    g.B = bl[l].Ain; g.sB = bl[l].sAin; g.ldb = bl[l].ldain; g.bMode = 1;
    g.C = G(bl[l].wseg); g.sC = P; g.ldc = s.seg[bl[l].wseg].ld;
    g.aAct = 1; g.bAct = bl[l].sAin != 0;
    g.epi = EPI_GRADW;
    // 3xFP16 split: the delta's tracked bound; the layer input's bound (Y: the slot of the most recent forward pass, which is
    // the one this backward pass differentiates)
    const int dsl[6] = {AMAX_DOUT, AMAX_D5, AMAX_D4, AMAX_DY, AMAX_D2, AMAX_D1};
    g.aSc = dsl[l];
    g.bSc = (l == 5) ? (int)SC_XTRAIN : (l == 2 ? (int)AMAX_Y : (int)SC_UNIT);
    g.cSc = -1;
    if (bl[l].bseg >= 0) { g.gbias = G(bl[l].bseg); g.sGb = P; }
    {
      const int ks = grad_ksplit(s, h->grid_hint, g, N);
      g.ksplit = ks;
      if (!tcm) {   // FP32 path: one gradient copy per segment, summed by the Adam phase
        s.seg_ksplit[bl[l].wseg] = ks;
        if (bl[l].bseg >= 0) s.seg_ksplit[bl[l].bseg] = ks;
        s.max_ksplit = std::max(s.max_ksplit, ks);
      }
    }
    if (ones) g.biasCol = bl[l].din;
    if (s.preconv && l == 5) { g.bImg = h->ximgM; g.sBImg = 0; g.bKbs = h->ximgM_kbs; }   // the training matrix, ones column included
    if (s.preconv && l != 5 && s.himg[4 - l]) {   // the layer input: H5, H4, Y, H2, H1
      const int k = 4 - l;
      g.bImg = s.himg[k]; g.sBImg = s.himg_stride[k]; g.bKbs = s.himg_kbs[k];
    }
    g.needLang = 1;
    add_maps(g);
    bw_idx[l] = ng;
    s.gemm[ng++] = g;
    bx_idx[l] = -1;
    if (bl[l].dIn) {
      // dIn[n,i] = sum_o dOut[n,o] * W[o,i]   (x sigmoid' of the layer input)
      GemmDesc x = make_gemm(s, N, bl[l].din, bl[l].dout, 1);
      x.A = bl[l].dOut; x.sA = bl[l].sD; x.lda = bl[l].ldd; x.aMode = 0;
      x.B = W(bl[l].wseg); x.sB = P; x.ldb = s.seg[bl[l].wseg].ld; x.bMode = 1;
      x.C = bl[l].dIn; x.sC = bl[l].sDin; x.ldc = bl[l].lddin;
      x.aAct = 1; x.cAct = 1;
      x.epi = bl[l].epi_x;
      x.aSc = dsl[l]; x.bSc = AMAX_W;
      x.cSc = (bl[l].epi_x == EPI_DSIG) ? dsl[l + 1] : (int)AMAX_DYIN;   // DY's own bound is derived by the BatchNorm backward
      if (s.preconv) {
        const WImg& wi = s.wi[wimg_index(bl[l].wseg)];
        x.bImg = s.wimg + wi.offM; x.sBImg = s.wimg_stride; x.bKbs = wi.kbsM;
      }
      if (bl[l].Hin) { x.aux = bl[l].Hin; x.sAux = bl[l].sDin; x.ldaux = bl[l].lddin; x.auxAct = 1; }
      x.needLang = 1;
      add_maps(x);
      bx_idx[l] = ng;
      s.gemm[ng++] = x;
    }
  }
  s.n_gemm = ng;
  // ---- phases
  auto set = [&](int pi, int kind, int g0, int g1, int arg, int rows, int need) {
    PhaseDesc& p = s.phase[pi];
    p.kind = kind; p.g0 = g0; p.g1 = g1; p.arg = arg; p.argRows = rows; p.needLang = need;
    p.sched_off = -1; p.sched_max = 0;
  };
  set(P_PRO, PH_PROLOGUE, -1, -1, 0, 0, 0);
  const int fstart[3] = {P_FWD1, P_FWD2, P_FWDT};
  for (int pass = 0; pass < 3; ++pass) {
    const int b = fwd_base[pass], p0 = fstart[pass];
    const int rows = (pass == 2) ? s.n_test : s.n_train;
    const int need = (pass == 0) ? 1 : 0;
    set(p0 + 0, PH_GEMM, b + 0, -1, 0, 0, need);
    set(p0 + 1, PH_GEMM, b + 1, -1, 0, 0, need);
    set(p0 + 2, PH_GEMM, b + 2, -1, 0, 0, need);
    set(p0 + 3, PH_BN_FWD, -1, -1, pass, rows, need);
    set(p0 + 4, PH_GEMM, b + 3, -1, 0, 0, need);
    set(p0 + 5, PH_GEMM, b + 4, -1, 0, 0, need);
    set(p0 + 6, PH_GEMM, b + 5, -1, 0, 0, need);
  }
  const int bstart[2] = {P_BWD1, P_BWD2};
  // tcgen05 path: the lagged order below unless BAE_TC_LAG=0 (A/B measurements); the FP32 path keeps the side-by-side order
  const bool lag = tcm && !(getenv("BAE_TC_LAG") && getenv("BAE_TC_LAG")[0] == '0');
  for (int pass = 0; pass < 2; ++pass) {
    const int p0 = bstart[pass];
    if (!lag) {
      // weight-gradient GEMM and delta GEMM of a layer side by side: both read the layer's output delta in the same phase
      set(p0 + 0, PH_GEMM, bw_idx[0], bx_idx[0], 0, 0, 1);
      set(p0 + 1, PH_GEMM, bw_idx[1], bx_idx[1], 0, 0, 1);
      set(p0 + 2, PH_GEMM, bw_idx[2], bx_idx[2], 0, 0, 1);
      set(p0 + 3, PH_BN_BWD, -1, -1, pass, s.n_train, 1);
      set(p0 + 4, PH_GEMM, bw_idx[3], bx_idx[3], 0, 0, 1);
      set(p0 + 5, PH_GEMM, bw_idx[4], bx_idx[4], 0, 0, 1);
      set(p0 + 6, PH_GEMM, bw_idx[5], -1, 0, 0, 1);
      continue;
    }
    // The weight-gradient GEMM of a layer runs ONE PHASE AFTER the delta GEMM of the same layer (neither depends on the other).
    // Measured at Madelon, 64 chains: backward phases -4.7% (2539 -> 2420 us per pass): the last phase no longer consists of
    // 4 long weight-gradient tiles per chain (256 items on 74 CTA pairs: a quarter of the last round idle) and the first one
    // is 1024 short delta tiles. The price is a second HBM read of every delta (+27 MB per chain and pass); HBM is far from
    // being the binding unit of these phases. (Persisting the split deltas as images for the weight-gradient GEMMs, which this
    // order allows, was measured too: their K blocks ran 12% faster without any conversion, the image writes - 33 MB per
    // chain-iteration - cost the delta GEMMs as much: not kept.)
    set(p0 + 0, PH_GEMM, bx_idx[0], -1, 0, 0, 1);                  // dH5
    set(p0 + 1, PH_GEMM, bw_idx[0], bx_idx[1], 0, 0, 1);           // dW6 | dH4
    set(p0 + 2, PH_GEMM, bw_idx[1], bx_idx[2], 0, 0, 1);           // dW5 | dY
    set(p0 + 3, PH_BN_BWD, -1, -1, pass, s.n_train, 1);
    set(p0 + 4, PH_GEMM, bw_idx[2], bx_idx[3], 0, 0, 1);           // dW4 | dH2 (reads dY behind the BatchNorm backward)
    set(p0 + 5, PH_GEMM, bw_idx[3], bx_idx[4], 0, 0, 1);           // dW3 | dH1
    set(p0 + 6, PH_GEMM, bw_idx[4], bw_idx[5], 0, 0, 1);           // dW2 | dW1
  }
  set(P_ADAM1, PH_ADAM1, -1, -1, 0, 0, 0);
  set(P_ADAM2, PH_ADAM2, -1, -1, 0, 0, 1);
  set(P_WAMAX0, PH_WAMAX, -1, -1, 1, 0, 0);   // arg 1: gated on "the previous iteration ended with an exchange round"
  set(P_EPI, PH_EPILOGUE, -1, -1, 0, 0, 0);
  set(P_SWAP + 0, PH_SWAP_DECIDE, -1, -1, 0, 0, 0);
  set(P_SWAP + 1, PH_SWAP_GATHER, -1, -1, 0, 0, 0);
  set(P_SWAP + 2, PH_SWAP_SCATTER, -1, -1, 0, 0, 0);
  s.n_phase = P_END;
}

// Longest-first (LPT) schedules of the GEMM phases of the tcgen05 path for launches over `nch` chains on `npairs` CTA
// pairs: a weight-gradient tile (all K segments, ~4 x 40 K blocks at Madelon) is several times longer than the other tiles of
// its phase, and round-robin dealing would leave the pairs up to one such tile apart. Cost model: K blocks + 6 per segment
// (accumulator drain and hand-off). Every pair walks its own list; the lists of a phase hold each item exactly once.
static int build_schedule(bae_handle* h, int nch, int npairs) {
  DevState& s = h->hs;
  std::vector<int> all;
  for (int p = 0; p < s.n_phase; ++p) {
    PhaseDesc& ph = s.phase[p];
    ph.sched_off = -1; ph.sched_max = 0;
    if (ph.kind != PH_GEMM) continue;
    std::vector<std::pair<int, int>> items;   // (cost, item index)
    const GemmDesc& ga = s.gemm[ph.g0];
    const int ta = ga.tm * ga.tn, tb = ph.g1 >= 0 ? s.gemm[ph.g1].tm * s.gemm[ph.g1].tn : 0;
    for (int c = 0; c < nch; ++c)
      for (int t = 0; t < ta + tb; ++t) {
        const GemmDesc& g = t < ta ? ga : s.gemm[ph.g1];
        items.emplace_back((g.K + 15) / 16 + 6 * g.ksplit, c * (ta + tb) + t);
      }
    std::stable_sort(items.begin(), items.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
    std::vector<long long> load(npairs, 0);
    std::vector<std::vector<int>> lists(npairs);
    for (const auto& it : items) {
      int best = 0;
      for (int q = 1; q < npairs; ++q)
        if (load[q] < load[best]) best = q;
      load[best] += it.first;
      lists[best].push_back(it.second);
    }
    size_t mx = 0;
    for (auto& l : lists) mx = std::max(mx, l.size());
    ph.sched_off = (int)all.size();
    ph.sched_max = (int)mx;
    for (auto& l : lists) {
      // items of one chain next to each other (operands shared in L2), longest first inside the list as dealt
      for (size_t k = 0; k < mx; ++k) all.push_back(k < l.size() ? l[k] : -1);
    }
  }
  if (all.empty()) return BAE_OK;
  void* q = nullptr;
  CU(cudaMalloc(&q, all.size() * sizeof(int)));
  h->allocs.push_back(q);
  CU(cudaMemcpyAsync(q, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  s.sched = reinterpret_cast<int*>(q);
  s.sched_chains = nch;
  s.sched_pairs = npairs;
  return BAE_OK;
}

static int upload_state(bae_handle* h) {
  CU(cudaMemcpyAsync(h->dS, &h->hs, sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
  return BAE_OK;
}

extern "C" {

const char* bae_last_error(void) { return g_err.c_str(); }
int bae_version(void) { return 100; }

int bae_create(const bae_config* cfg, int device, bae_handle** out) {
  if (!cfg || !out) return fail(BAE_ERR_INVALID, "null argument");
  for (int i = 0; i < 4; ++i)
    if (cfg->dims[i] < 1) return fail(BAE_ERR_INVALID, "dims must be >= 1");
  if (cfg->n_train < 2 || cfg->n_test < 2) return fail(BAE_ERR_INVALID, "need at least 2 rows (BatchNorm batch statistics)");
  if (cfg->n_chains < 1 || cfg->n_rungs < 1 || cfg->n_chains % cfg->n_rungs != 0)
    return fail(BAE_ERR_INVALID, "n_chains must be a positive multiple of n_rungs");
  if (cfg->swap_interval < 1 || cfg->samples < 1) return fail(BAE_ERR_INVALID, "swap_interval and samples must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(BAE_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(BAE_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                       ", this library is built for sm_100a only");
  CU(cudaSetDevice(device));
  bae_handle* h = new bae_handle();
  // every failure path below releases the handle, its stream / events and all device allocations made so far
  struct Guard {
    bae_handle* h;
    ~Guard() { if (h) bae_destroy(h); }
  } guard{h};
  h->cfg = *cfg;
  h->device = device;
  h->phase_launch = getenv("BAE_PHASE_LAUNCH") != nullptr && getenv("BAE_PHASE_LAUNCH")[0] == '1';
  h->phase_launch_env = getenv("BAE_PHASE_LAUNCH") != nullptr;
  h->use_graph = !(getenv("BAE_NO_GRAPH") != nullptr && getenv("BAE_NO_GRAPH")[0] == '1');
  h->phase_times = getenv("BAE_PHASE_TIMES") != nullptr && getenv("BAE_PHASE_TIMES")[0] == '1';
  if (h->phase_times) h->use_graph = false;
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  DevState& s = h->hs;
  memset(&s, 0, sizeof(s));
  s.d0 = cfg->dims[0]; s.d1 = cfg->dims[1]; s.d2 = cfg->dims[2]; s.d3 = cfg->dims[3];
  {
    // tcgen05 path only where the layers are real dense contractions (Madelon-like widths); narrow layers
    // (Coil-2000's 50-85, Swiss Roll's 2-10) stay on the FP32 CUDA-core tiles.
    const int dmin = std::min(std::min(s.d0, s.d1), std::min(s.d2, s.d3));
    const bool rows_ok = cfg->n_train >= 128 && cfg->n_test >= 128;
    // automatic choice: wide layers (Madelon and wider) always; mid widths (Coil-2000's 50-85: tiles of 64-96 columns, measured
    // 2x the FP32 tiles at 64 chains, slower at 8 where a launch per phase costs more than the phase) when many chains are resident
    const bool eligible = rows_ok && (dmin >= 128 || (dmin >= 48 && cfg->n_chains >= 32));
    if (cfg->gemm_path == 2 && !(rows_ok && dmin >= 48))
      return fail(BAE_ERR_INVALID, "gemm_path=2 (tcgen05) needs every layer width >= 48 and >= 128 rows");
    s.tc = (cfg->gemm_path == 2 || (cfg->gemm_path == 0 && eligible)) ? 1 : 0;
    if (s.tc && !get_encode()) return fail(BAE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  }
  // row pitches: multiple of 4 floats; the tcgen05 path reserves one extra column (ones column, bias gradients)
  const int extra = s.tc ? 1 : 0;
  s.ld0 = round_up(s.d0 + extra, 4); s.ld1 = round_up(s.d1 + extra, 4); s.ld2 = round_up(s.d2 + extra, 4);
  s.ld3 = round_up(s.d3 + extra, 4);
  s.n_train = cfg->n_train; s.n_test = cfg->n_test; s.n_max = std::max(cfg->n_train, cfg->n_test);
  s.C = cfg->n_chains; s.C_alloc = s.C + 1; s.n_rungs = cfg->n_rungs; s.chain_id_base = cfg->chain_id_base;
  s.samples = cfg->samples; s.pt_step = cfg->pt_switch_step; s.swap_interval = cfg->swap_interval;
  s.use_langevin = cfg->use_langevin;
  if (cfg->drift_mode < 0 || cfg->drift_mode > 2) return fail(BAE_ERR_INVALID, "drift_mode must be 0 (Adam), 1 (Adam, SGD after pt) or 2 (MALA)");
  s.drift_mode = cfg->drift_mode;
  s.lr = cfg->lr; s.step_w = cfg->step_w; s.step_eta = cfg->step_eta; s.l_prob = cfg->l_prob;
  s.sigma_sq = cfg->sigma_squared; s.nu1 = cfg->nu_1; s.nu2 = cfg->nu_2;
  s.seed_lo = (unsigned)(cfg->seed & 0xffffffffull); s.seed_hi = (unsigned)(cfg->seed >> 32);
  build_layout(s);
  const size_t CA = s.C_alloc, P = s.Pp;
  CU(dalloc(h, &s.theta, CA * P)); CU(dalloc(h, &s.m, CA * P)); CU(dalloc(h, &s.v, CA * P));
  CU(dalloc(h, &s.wcur, CA * P));
  const bool tf32_env = getenv("BAE_TC_TF32") && getenv("BAE_TC_TF32")[0] == '1';
  // (the 3xFP16 engine belongs to the per-phase step program: see below where s.f16 is set)
  const bool f16_early = s.tc && (!h->phase_launch_env || h->phase_launch) && !tf32_env;
  s.preconv = (f16_early && !(getenv("BAE_TC_PRECONV") && getenv("BAE_TC_PRECONV")[0] == '0')) ? 1 : 0;
  if (s.preconv) {
    // pre-split images: per chain the six weight matrices (K-major for the forward GEMMs, MN-major for the delta GEMMs of all
    // layers but the first); once per GPU the data matrices
    long long off = 0;
    s.wimg_items = 0;
    for (int i = 0; i < 6; ++i) {
      const Seg& sd = s.seg[kWimgSegs[i]];
      WImg& wi = s.wi[i];
      wi.seg = kWimgSegs[i];
      wi.items = ((sd.rows + 7) / 8) * ((((sd.cols + 7) / 8) + 63) / 64);
      s.wimg_items += wi.items;
      wi.offK = off; wi.kbsK = img_kbs(sd.rows); off += (long long)img_bytes(sd.rows, sd.cols);
      if (i > 0) { wi.offM = off; wi.kbsM = img_kbs(sd.cols); off += (long long)img_bytes(sd.cols, sd.rows); }
      else { wi.offM = -1; wi.kbsM = 0; }
    }
    s.wimg_stride = off;
    CU(dalloc(h, &s.wimg, CA * (size_t)off));
    const int nx[2] = {cfg->n_train, cfg->n_test};
    for (int i = 0; i < 2; ++i) {
      h->ximgK_kbs[i] = img_kbs(nx[i]);
      CU(dalloc(h, &h->ximgK[i], img_bytes(nx[i], s.d0)));
    }
    h->ximgM_kbs = img_kbs(s.d0 + 1);
    CU(dalloc(h, &h->ximgM, img_bytes(s.d0 + 1, cfg->n_train)));
  }
  for (int i = 0; i < kNumSegs; ++i) s.seg_ksplit[i] = 1;
  s.max_ksplit = 1;
  {
    const int dd[7] = {s.d0, s.d1, s.d2, s.d3, s.d2, s.d1, s.d0};
    for (int l = 0; l < 6 && !s.tc; ++l)
      s.max_ksplit = std::max(s.max_ksplit, grad_ksplit(s, h->grid_hint, make_gemm(s, dd[l + 1], dd[l], s.n_train, 1), s.n_train));
  }
  CU(dalloc(h, &s.grad, (size_t)s.max_ksplit * CA * P));   // K-split copies of the gradient (copy 0 is the result)
  CU(dalloc(h, &s.wb, CA * s.nbuf)); CU(dalloc(h, &s.wb_tmp, CA * s.nbuf));
  CU(dalloc(h, &s.bn_mu, 3 * CA * s.ld3)); CU(dalloc(h, &s.bn_var, 3 * CA * s.ld3)); CU(dalloc(h, &s.bn_inv, CA * s.ld3));
  CU(dalloc(h, &s.Xtrain, (size_t)s.n_train * s.ld0)); CU(dalloc(h, &s.Xtest, (size_t)s.n_test * s.ld0));
  const size_t nm = s.n_max;
  {
    // chain groups: activation / delta buffers exist for G chains at a time (G = C unless they do not fit)
    size_t per_chain = nm * (size_t)(s.ld0 + 4 * s.ld1 + 4 * s.ld2 + 3 * s.ld3) * sizeof(float);   // H1 H2 Z Y H4 H5 + DOUT D5 D4 DY D2 D1
    if (s.preconv && !(getenv("BAE_TC_HIMG") && getenv("BAE_TC_HIMG")[0] == '0'))
      per_chain += 2 * (img_bytes(s.d1 + 1, s.n_train) + img_bytes(s.d2 + 1, s.n_train)) + img_bytes(s.d3 + 1, s.n_train);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    int G = s.C;
    if (cfg->act_group > 0) {
      G = std::min(s.C, cfg->act_group);
    } else {
      const size_t budget = (size_t)(0.70 * (double)free_b);
      if ((size_t)s.C * per_chain > budget) G = (int)std::max<size_t>(1, budget / per_chain);
    }
    s.G = std::max(1, std::min(G, s.C));
  }
  const size_t CA_act = s.G;   // the parity probe's scratch chain runs alone and uses slot 0
  CU(dalloc(h, &s.H1, CA_act * nm * s.ld1)); CU(dalloc(h, &s.H2, CA_act * nm * s.ld2)); CU(dalloc(h, &s.Z, CA_act * nm * s.ld3));
  CU(dalloc(h, &s.Y, CA_act * nm * s.ld3)); CU(dalloc(h, &s.H4, CA_act * nm * s.ld2)); CU(dalloc(h, &s.H5, CA_act * nm * s.ld1));
  CU(dalloc(h, &s.DOUT, CA_act * nm * s.ld0)); CU(dalloc(h, &s.D5, CA_act * nm * s.ld1)); CU(dalloc(h, &s.D4, CA_act * nm * s.ld2));
  CU(dalloc(h, &s.DY, CA_act * nm * s.ld3)); CU(dalloc(h, &s.D2, CA_act * nm * s.ld2)); CU(dalloc(h, &s.D1, CA_act * nm * s.ld1));
  if (s.preconv && !(getenv("BAE_TC_HIMG") && getenv("BAE_TC_HIMG")[0] == '0')) {
    // MN-major images of the activations (B operands of the weight-gradient GEMMs), written by the converter warps of the forward
    // GEMM that reads the activation as its A operand
    const int hd[5] = {s.d1, s.d2, s.d3, s.d2, s.d1};   // H1, H2, Y, H4, H5
    for (int i = 0; i < 5; ++i) {
      // Y (BatchNorm output, scale tracked per chain): when no forward K block reaches its ones column (d3 a multiple of 16) the
      // column would have to be static, which a per-chain scale does not allow - that one operand then stays with the converters
      if (i == 2 && hd[i] % 16 == 0) continue;
      s.himg_kbs[i] = img_kbs(hd[i] + 1);
      s.himg_stride[i] = (long long)img_bytes(hd[i] + 1, s.n_train);
      CU(dalloc(h, &s.himg[i], CA_act * (size_t)s.himg_stride[i]));
      // (static content of the ones column of a sigmoid output - static scale 2^13 - where no forward K block reaches it)
      if (hd[i] % 16 == 0) CU(launch_himg_ones(s.himg[i], s.himg_stride[i], s.himg_kbs[i], hd[i], s.n_train, (int)CA_act, h->stream));
    }
  }
  {
    const int per_tile = s.tc ? 2 : 1;   // tcgen05 path: each CTA of a pair writes the partial of its 128 rows
    GemmDesc t = make_gemm(s, s.n_max, s.d0, s.d1);
    s.loss_stride = per_tile * t.tm * t.tn;
    GemmDesc t2 = make_gemm(s, std::min(s.n_train, s.n_test), s.d0, s.d1);
    s.loss_stride = std::max(s.loss_stride, per_tile * t2.tm * t2.tn);
  }
  if (s.tc) {
    // ones column (index d) of every activation that is the B operand of a weight-gradient GEMM
    std::vector<float> ones(CA_act * nm, 1.0f);
    struct OC { float* p; int ld, d; };
    const OC oc[5] = {{s.H1, s.ld1, s.d1}, {s.H2, s.ld2, s.d2}, {s.Y, s.ld3, s.d3}, {s.H4, s.ld2, s.d2}, {s.H5, s.ld1, s.d1}};
    for (const OC& o : oc)
      CU(cudaMemcpy2DAsync(o.p + o.d, (size_t)o.ld * 4, ones.data(), 4, 4, CA_act * nm, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  CU(dalloc(h, &s.loss_part, 3 * CA * (size_t)s.loss_stride * 2));
  CU(dalloc(h, &s.adam_part, CA * (size_t)s.n_adam_items * 3));
  ChainArrays& c = s.ch;
  CU(dalloc(h, &c.eta, CA)); CU(dalloc(h, &c.lik, CA)); CU(dalloc(h, &c.prior_cur, CA)); CU(dalloc(h, &c.last_sse, CA)); CU(dalloc(h, &c.prop_sse, CA));
  CU(dalloc(h, &c.mse_tr, CA)); CU(dalloc(h, &c.mse_te, CA)); CU(dalloc(h, &c.T, CA)); CU(dalloc(h, &c.adapt, CA));
  CU(dalloc(h, &c.adam_t, CA)); CU(dalloc(h, &c.alias, CA)); CU(dalloc(h, &c.init_count, CA)); CU(dalloc(h, &c.nacc, CA));
  CU(dalloc(h, &c.nlang, CA)); CU(dalloc(h, &c.step, CA)); CU(dalloc(h, &c.is_lang, CA));
  CU(dalloc(h, &c.lx, CA)); CU(dalloc(h, &c.eta_noise, CA)); CU(dalloc(h, &c.u, CA));
  CU(dalloc(h, &c.ss1, CA)); CU(dalloc(h, &c.sq1, CA)); CU(dalloc(h, &c.ss2, CA)); CU(dalloc(h, &c.sq2, CA));
  CU(dalloc(h, &c.plain, CA)); CU(dalloc(h, &c.drift_a, CA)); CU(dalloc(h, &c.drift_b, CA));
  CU(dalloc(h, &c.gid, CA));
  {
    std::vector<int> gid(CA);
    for (size_t i = 0; i < CA; ++i) gid[i] = cfg->chain_id_base + (int)i;
    CU(cudaMemcpyAsync(c.gid, gid.data(), sizeof(int) * CA, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  s.shard_rank = 0; s.shard_world = 1;
  h->n_sms = prop.multiProcessorCount;
  {
    const Seg& last = s.seg[kNumSegs - 1];
    h->p_used = last.pad_off + round_up(last.rows * last.ld, 32);
    const bool off = getenv("BAE_NO_NARROW") && getenv("BAE_NO_NARROW")[0] == '1';
    h->narrow = !off && !s.tc && cfg->gemm_path != 1 && s.drift_mode == 0 && narrow_supported(cfg->dims, cfg->n_train, cfg->n_test, h->p_used);
    // few chains: more CTAs per chain (the rows of the data set are cut over the cluster); many chains: one CTA each
    h->narrow_cluster = s.C <= 16 ? 8 : (s.C <= 32 ? 4 : (s.C <= 64 ? 2 : 1));
    if (const char* e = getenv("BAE_NARROW_CLUSTER")) {
      const int v = atoi(e);
      if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) h->narrow_cluster = v;
    }
  }
  CU(dalloc(h, &s.traces, (size_t)s.C * s.samples));
  CU(dalloc(h, &s.swap_src, CA)); CU(dalloc(h, &s.swap_flag, CA)); CU(dalloc(h, &s.swap_tmp, CA * 2));
  CU(dalloc(h, &s.num_swap, 1)); CU(dalloc(h, &s.total_swap, 1));
  CU(dalloc(h, &h->stage_ref, (size_t)s.w_size + 8)); CU(dalloc(h, &h->stage_noise, (size_t)s.w_size + 8));
  CU(dalloc(h, &h->stage_d, 64));
  CU(dalloc(h, &h->stage_sc, 1));
  if (getenv("BAE_TC_TRACE") && getenv("BAE_TC_TRACE")[0] == '1') CU(dalloc(h, &s.dbg, 4096));
  s.dbg_flags = getenv("BAE_TC_DBGFLAGS") ? atoi(getenv("BAE_TC_DBGFLAGS")) : 0;
#ifndef BAE_TC_DEBUG
  // the timing-experiment branches (results are wrong when one is set) are compiled only with -DBAE_TC_DEBUG
  if (s.dbg_flags != 0)
    return fail(BAE_ERR_INVALID, "BAE_TC_DBGFLAGS is set, but this library was built without -DBAE_TC_DEBUG: refusing to sample");
#endif
  for (const char* knob : {"BAE_TC_KSPLIT_ROWS", "BAE_TC_MAXBN", "BAE_PHASE_LAUNCH", "BAE_NO_GRAPH", "BAE_TC_TRACE", "BAE_TC_NO_SCHED", "BAE_NO_NARROW", "BAE_NARROW_CLUSTER", "BAE_TC_TF32", "BAE_TC_PRECONV", "BAE_TC_HIMG", "BAE_TC_LAG"})
    if (getenv(knob)) fprintf(stderr, "[bae_b200] note: tuning variable %s=%s is in effect for this handle\n", knob, getenv(knob));
  s.dbg_phase = getenv("BAE_TC_TRACE_PHASE") ? atoi(getenv("BAE_TC_TRACE_PHASE")) : -1;
  {
    void* q = nullptr;
    CU(cudaMalloc(&q, sizeof(DevState)));
    h->allocs.push_back(q);
    h->dS = reinterpret_cast<DevState*>(q);
  }
  if (s.tc && !h->phase_launch_env) h->phase_launch = true;   // (reasons below, where the grids are sized)
  {
    // 3xFP16 split (kind::f16, half the tensor-pipe time of 3xTF32 at the same 11-bit operand precision): the default of the
    // per-phase step program; BAE_TC_TF32=1 keeps the 3xTF32 engine, and the persistent kernel always uses it (the amax
    // slots are zeroed by launches between phases)
    s.f16 = (s.tc && h->phase_launch && !tf32_env) ? 1 : 0;
    if (s.preconv && !s.f16) return fail(BAE_ERR_STATE, "internal: operand images without the 3xFP16 engine");
    CU(dalloc(h, &s.amax, (size_t)kAmaxSlots * CA));
    s.xexp[0] = s.xexp[1] = kF16TopExp;
  }
  build_program(h);
  if (s.tc) {
    if (!h->tmap_error.empty()) return fail(BAE_ERR_CUDA, h->tmap_error);
    void* q = nullptr;
    CU(cudaMalloc(&q, h->tmaps_host.size() * sizeof(CUtensorMap)));
    h->allocs.push_back(q);
    CU(cudaMemcpyAsync(q, h->tmaps_host.data(), h->tmaps_host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, h->stream));
    s.tmaps = q;
  }
  int rc;
  const int bps = max_coop_blocks_per_sm(s.tc);
  if (bps < 1) return fail(BAE_ERR_CUDA, "step kernel cannot be made resident (occupancy 0)");
  h->grid = prop.multiProcessorCount * bps;
  if (s.tc) {   // CTA pairs (cluster dimension 2): as many pairs as can be co-resident
    h->grid = max_coop_grid_tc();
    if (h->grid < 2) return fail(BAE_ERR_CUDA, "the tcgen05 step kernel (CTA pairs) cannot be made resident");
    for (int k = 0; k <= PH_WAMAX; ++k) {
      const int abs = aux_blocks_per_sm(k);
      if (abs < 1) return fail(BAE_ERR_CUDA, "bae_aux_kernel cannot be made resident");
      h->grid_aux[k] = prop.multiProcessorCount * abs;
    }
    // Measured (Madelon, 64 chains): BatchNorm / Adam phases run 1.5-1.8x faster as their own many-CTA launches than on the two
    // epilogue warpgroups of the persistent kernel, and a launch costs no more than a grid barrier.
    if (!h->phase_launch_env) h->phase_launch = true;
    if (s.drift_mode != 0 && !h->phase_launch)
      return fail(BAE_ERR_INVALID, "drift_mode 1 / 2 on the tcgen05 path needs the per-phase step program (unset BAE_PHASE_LAUNCH=0)");
    if (!(getenv("BAE_TC_NO_SCHED") && getenv("BAE_TC_NO_SCHED")[0] == '1'))
      if ((rc = build_schedule(h, s.G, h->grid / 2))) return rc;
  }
  if ((rc = upload_state(h))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  guard.h = nullptr;
  *out = h;
  return BAE_OK;
}

int bae_destroy(bae_handle* h) {
  if (!h) return BAE_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->xsrc_host) cudaFreeHost(h->xsrc_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& pe : h->pt_events) cudaEventDestroy(pe.second);
  cudaStreamDestroy(h->stream);
  delete h;
  return BAE_OK;
}

int bae_w_size(const bae_handle* h) { return h ? h->hs.w_size : BAE_ERR_INVALID; }
int bae_n_trainable(const bae_handle* h) { return h ? h->hs.w_size - h->hs.nbuf : BAE_ERR_INVALID; }
int bae_gemm_path(const bae_handle* h) { return h ? (h->hs.tc ? 2 : 1) : BAE_ERR_INVALID; }
int bae_mma_kind(const bae_handle* h) { return h ? (h->hs.tc ? (h->hs.f16 ? 2 : 1) : 0) : BAE_ERR_INVALID; }
int bae_step_kernel(const bae_handle* h) { return h ? (h->narrow ? 3 : (h->hs.tc ? 2 : 1)) : BAE_ERR_INVALID; }
void* bae_stream(bae_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t bae_launch_count(const bae_handle* h) { return h ? h->launches : 0; }

int bae_synchronize(bae_handle* h) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

int bae_upload_data(bae_handle* h, const float* train, const float* test) {
  if (!h || !train || !test) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpy2DAsync(s.Xtrain, (size_t)s.ld0 * 4, train, (size_t)s.d0 * 4, (size_t)s.d0 * 4, s.n_train,
                       cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpy2DAsync(s.Xtest, (size_t)s.ld0 * 4, test, (size_t)s.d0 * 4, (size_t)s.d0 * 4, s.n_test,
                       cudaMemcpyHostToDevice, h->stream));
  if (s.tc) {
    // ones column at index d0: the weight-gradient GEMM of encode.0 reads it to produce the bias gradient
    if (h->ones_col.size() < (size_t)s.n_max) h->ones_col.assign(s.n_max, 1.0f);
    CU(cudaMemcpy2DAsync(s.Xtrain + s.d0, (size_t)s.ld0 * 4, h->ones_col.data(), 4, 4, s.n_train, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpy2DAsync(s.Xtest + s.d0, (size_t)s.ld0 * 4, h->ones_col.data(), 4, 4, s.n_test, cudaMemcpyHostToDevice, h->stream));
  }
  if (s.f16) {
    // bounds of the two data matrices for the FP16 operand scales (their ones column included)
    // (sixteen independent running maxima: the scan is on the caller's critical path of every upload; `x - x` is 0 for finite
    //  values and NaN otherwise, so one sum finds non-finite data without a test per element)
    auto absmax = [](const float* p, size_t n, float& bad) {
      float m[16], z[16];
      for (int k = 0; k < 16; ++k) { m[k] = 1.0f; z[k] = 0.f; }
      size_t i = 0;
      for (; i + 16 <= n; i += 16)
        for (int k = 0; k < 16; ++k) {
          const float v = p[i + k], a = std::fabs(v);
          m[k] = a > m[k] ? a : m[k];
          z[k] += v - v;
        }
      for (; i < n; ++i) { const float v = p[i], a = std::fabs(v); m[0] = a > m[0] ? a : m[0]; z[0] += v - v; }
      float r = m[0];
      for (int k = 1; k < 16; ++k) { r = m[k] > r ? m[k] : r; z[0] += z[k]; }
      bad = z[0];
      return r;
    };
    float bad[2];
    float mx[2] = {absmax(train, (size_t)s.n_train * s.d0, bad[0]), absmax(test, (size_t)s.n_test * s.d0, bad[1])};
    if (!(bad[0] == 0.f) || !(bad[1] == 0.f) || !std::isfinite(mx[0]) || !std::isfinite(mx[1]))
      return fail(BAE_ERR_INVALID, "data contains non-finite values");
    h->hs.xexp[0] = f16_scale_exp(mx[0]);
    h->hs.xexp[1] = f16_scale_exp(mx[1]);
    int rc = upload_state(h);
    if (rc) return rc;
    if (s.preconv) {
      // pre-split images of the matrices just uploaded (stream-ordered behind the copies above)
      CU(launch_ximg(s.Xtrain, s.ld0, s.n_train, s.d0, s.d0 + 1, h->hs.xexp[0], h->ximgK[0], h->ximgK_kbs[0], h->ximgM, h->ximgM_kbs, h->stream));
      CU(launch_ximg(s.Xtest, s.ld0, s.n_test, s.d0, 0, h->hs.xexp[1], h->ximgK[1], h->ximgK_kbs[1], nullptr, 0, h->stream));
      h->launches += 2;
    }
  }
  h->data_loaded = true;
  return BAE_OK;
}

int bae_set_ladder(bae_handle* h, const double* T) {
  if (!h || !T) return fail(BAE_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->hs.ch.T, T, sizeof(double) * h->hs.C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->hs.ch.adapt, T, sizeof(double) * h->hs.C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

}  // extern "C"

// ---- helpers -----------------------------------------------------------------------------------
static int put_vec(bae_handle* h, const float* host_ref, float* dev_pad_chain) {
  const DevState& s = h->hs;
  CU(cudaMemcpyAsync(h->stage_ref, host_ref, sizeof(float) * s.w_size, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(dev_pad_chain, 0, sizeof(float) * s.Pp, h->stream));
  CU(launch_pack(h->dS, h->stage_ref, dev_pad_chain, h->stream));
  h->launches++;
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

static int get_vec(bae_handle* h, const float* dev_pad_chain, float* host_ref) {
  const DevState& s = h->hs;
  CU(launch_unpack(h->dS, dev_pad_chain, h->stage_ref, h->stream));
  h->launches++;
  CU(cudaMemcpyAsync(host_ref, h->stage_ref, sizeof(float) * s.w_size, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

template <typename T>
static int put1(bae_handle* h, T* dev_arr, int idx, T v) {
  CU(cudaMemcpyAsync(dev_arr + idx, &v, sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}
template <typename T>
static int get1(bae_handle* h, const T* dev_arr, int idx, T* v) {
  CU(cudaMemcpyAsync(v, dev_arr + idx, sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

// One phase as its own launch: GEMM phases (and everything on the FP32 path) through the phase kernel of the path, the
// non-GEMM phases of the tcgen05 path through bae_aux_kernel.
// 3xFP16 split: the amax slots a phase range produces are zeroed by one small launch in front of its first phase.
static unsigned amax_zero_mask(int p) {
  if (p == P_WAMAX0 || p == P_ADAM1) return 1u << AMAX_W;
  if (p == P_FWD1 || p == P_FWD2 || p == P_FWDT) return (1u << AMAX_Y) | (1u << AMAX_DOUT);
  if (p == P_BWD1 || p == P_BWD2) return (1u << AMAX_D5) | (1u << AMAX_D4) | (1u << AMAX_DY) | (1u << AMAX_DYIN) | (1u << AMAX_D2) | (1u << AMAX_D1);
  return 0u;
}

static cudaError_t launch_one(bae_handle* h, int p, int c_begin, int c_end, int force_swap) {
  if (h->hs.phase[p].kind == PH_WAMAX && !h->hs.f16) return cudaSuccess;
  if (h->hs.f16) {
    const unsigned mask = amax_zero_mask(p);
    if (mask) {
      const cudaError_t e = launch_amax_zero(h->dS, mask, c_begin, c_end, (p == P_WAMAX0 && !force_swap) ? 1 : 0, h->stream);
      if (e != cudaSuccess) return e;
      h->n_emit++;
    }
  }
  h->n_emit++;
  if (h->hs.preconv && (p == P_WAMAX0 || p == P_ADAM1)) {
    // the phase that folds the weight bounds, then the images of the weights it bounds (same gate)
    const cudaError_t e = launch_aux(h->dS, h->hs.phase[p].kind, h->grid_aux[h->hs.phase[p].kind], p, c_begin, c_end, force_swap, h->stream);
    if (e != cudaSuccess) return e;
    h->n_emit++;
    return launch_wconv(h->dS, h->n_sms * 8, c_begin, c_end, (p == P_WAMAX0 && !force_swap) ? 1 : 0, h->stream);
  }
  if (h->hs.tc && h->hs.phase[p].kind != PH_GEMM)
    return launch_aux(h->dS, h->hs.phase[p].kind, h->grid_aux[h->hs.phase[p].kind], p, c_begin, c_end, force_swap, h->stream);
  const PhaseDesc& ph = h->hs.phase[p];
  const int ea = ph.kind == PH_GEMM ? h->hs.gemm[ph.g0].epi : -1;
  const int eb = (ph.kind == PH_GEMM && ph.g1 >= 0) ? h->hs.gemm[ph.g1].epi : -1;
  // what every GEMM of the phase does with operand images (bae_tc.cuh: FrontMode; 0 = decided per GEMM at run time)
  int mode = 0;
  if (ph.kind == PH_GEMM && !h->hs.preconv && !h->hs.dbg) mode = 4;   // FM_NONE: no operand images anywhere
  if (ph.kind == PH_GEMM && h->hs.preconv && !h->hs.dbg) {   // (BAE_TC_TRACE=1: the timeline trace lives in the generic kernels)
    const GemmDesc& ga = h->hs.gemm[ph.g0];
    const GemmDesc* gb = ph.g1 >= 0 ? &h->hs.gemm[ph.g1] : nullptr;
    const bool allB = ga.bImg && (!gb || gb->bImg), anyA = ga.aImg || (gb && gb->aImg), anyOut = ga.aOut || (gb && gb->aOut);
    if (allB && !anyA && !anyOut) mode = 1;                       // FM_B
    else if (allB && !gb && ga.aImg && !ga.aOut) mode = 2;        // FM_AB
    else if (allB && !gb && !ga.aImg && ga.aOut) mode = 3;        // FM_B_OUT
  }
  return launch_phase(h->dS, h->hs.tc, h->grid, p, c_begin, c_end, force_swap, ea, eb, mode, h->stream);
}

static int run_phases_plain(bae_handle* h, int p_begin, int p_end, int c_begin, int c_end) {
  for (int p = p_begin; p < p_end; ++p) {
    const int64_t n0 = h->n_emit;
    CU(launch_one(h, p, c_begin, c_end, 1));
    h->launches += h->n_emit - n0;
  }
  return BAE_OK;
}

// forward of slot `pass` (0: train via FWD1 phases, 2: test via FWDT phases) for one chain; returns SSE and sum(d)
static int forward_one(bae_handle* h, int chain, int pass, double* sse, double* sumd) {
  const DevState& s = h->hs;
  int rc = put1<int>(h, s.ch.is_lang, chain, 1);
  if (rc) return rc;
  const int p0 = (pass == 0) ? P_FWD1 : P_FWDT;
  rc = run_phases_plain(h, P_WAMAX0, P_WAMAX0 + 1, chain, chain + 1);   // 3xFP16 split: bounds of the weights just written
  if (rc) return rc;
  rc = run_phases_plain(h, p0, p0 + 7, chain, chain + 1);
  if (rc) return rc;
  std::vector<double> part((size_t)s.loss_stride * 2);
  CU(cudaMemcpyAsync(part.data(), s.loss_part + ((size_t)pass * s.C_alloc + chain) * s.loss_stride * 2,
                     sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  double a = 0, b = 0;
  for (int i = 0; i < s.loss_stride; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
  *sse = a;
  if (sumd) *sumd = b;
  return BAE_OK;
}

static double loglik_h(double sse, double n, double tau_sq, double temp) {
  return (-(n * 0.5) * std::log(2.0 * M_PI * tau_sq) - 0.5 * tau_sq * sse) / temp;
}

extern "C" {

int bae_init_chain(bae_handle* h, int chain, const float* theta_init, const double* w0) {
  if (!h || !theta_init || !w0) return fail(BAE_ERR_INVALID, "null argument");
  if (!h->data_loaded) return fail(BAE_ERR_STATE, "upload data before initialising chains");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  float* th = s.theta + (size_t)chain * s.Pp;
  int rc = put_vec(h, theta_init, th);
  if (rc) return rc;
  // eta0 = log var(pred_train - train) with the torch-default-initialised model   MAIN:381-385
  double sse = 0, sd = 0;
  rc = forward_one(h, chain, 0, &sse, &sd);
  if (rc) return rc;
  const double n = (double)s.n_train * s.d0;
  const double var = sse / n - (sd / n) * (sd / n);
  const double eta = std::log(var);
  const double tau = std::exp(eta);
  // w_proposal = dictfromlist(randn) -> fp32; prior on the UNquantised float64 vector   MAIN:389-399
  std::vector<float> wq(s.w_size);
  double sumsq = 0;
  for (int i = 0; i < s.w_size; ++i) { wq[i] = (float)w0[i]; sumsq += w0[i] * w0[i]; }
  const double prior = -((double)s.w_size * 0.5) * std::log((double)s.sigma_sq) - sumsq / (2.0 * (double)s.sigma_sq) -
                       (1.0 + (double)s.nu1) * std::log(tau) - ((double)s.nu2 / tau);
  const int nbt_ref = s.seg[SEG_BN_NBT].ref_off, rm_ref = s.seg[SEG_BN_RM].ref_off, rv_ref = s.seg[SEG_BN_RV].ref_off;
  std::vector<float> loaded = wq;
  loaded[nbt_ref] = std::trunc(wq[nbt_ref]);   // load_state_dict into the int64 buffer
  rc = put_vec(h, loaded.data(), th);
  if (rc) return rc;
  rc = forward_one(h, chain, 0, &sse, nullptr);                                       // MAIN:401
  if (rc) return rc;
  double T = 1.0;
  rc = get1<double>(h, s.ch.T, chain, &T);
  if (rc) return rc;
  const double lik = loglik_h(sse, n, tau, T);
  // the forward updated the live model's running statistics
  std::vector<float> mu(s.ld3), va(s.ld3);
  CU(cudaMemcpyAsync(mu.data(), s.bn_mu + ((size_t)0 * s.C_alloc + chain) * s.ld3, sizeof(float) * s.ld3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(va.data(), s.bn_var + ((size_t)0 * s.C_alloc + chain) * s.ld3, sizeof(float) * s.ld3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const float ub = (float)s.n_train / (float)(s.n_train - 1);
  for (int c = 0; c < s.d3; ++c) {
    loaded[rm_ref + c] = 0.9f * loaded[rm_ref + c] + 0.1f * mu[c];
    loaded[rv_ref + c] = 0.9f * loaded[rv_ref + c] + 0.1f * va[c] * ub;
  }
  loaded[nbt_ref] += 1.0f;
  rc = put_vec(h, loaded.data(), th);
  if (rc) return rc;
  CU(cudaMemsetAsync(s.m + (size_t)chain * s.Pp, 0, sizeof(float) * s.Pp, h->stream));
  CU(cudaMemsetAsync(s.v + (size_t)chain * s.Pp, 0, sizeof(float) * s.Pp, h->stream));
  bae_chain_scalars sc;
  memset(&sc, 0, sizeof(sc));
  sc.eta = eta; sc.likelihood = lik; sc.prior_current = prior; sc.last_sse_train = sse; sc.prop_sse_train = sse;
  sc.temperature = T; sc.adapttemp = T; sc.w_is_alias = 1;
  return bae_set_state(h, chain, nullptr, nullptr, nullptr, nullptr, &sc);   // (sets the host's iteration mirror to 0)
}

int bae_get_state(bae_handle* h, int chain, float* theta, float* w, float* m, float* v, bae_chain_scalars* sc) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  int rc;
  const size_t o = (size_t)chain * s.Pp;
  if (theta && (rc = get_vec(h, s.theta + o, theta))) return rc;
  if (m && (rc = get_vec(h, s.m + o, m))) return rc;
  if (v && (rc = get_vec(h, s.v + o, v))) return rc;
  if (w) {
    // `w` = trainables of the live model (they coincide at every step boundary) + the detached buffers
    if ((rc = get_vec(h, s.theta + o, w))) return rc;
    int alias = 1;
    if ((rc = get1<int>(h, s.ch.alias, chain, &alias))) return rc;
    if (!alias) {
      std::vector<float> wb(s.nbuf);
      CU(cudaMemcpyAsync(wb.data(), s.wb + (size_t)chain * s.nbuf, sizeof(float) * s.nbuf, cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
      for (int c = 0; c < s.d3; ++c) {
        w[s.seg[SEG_BN_RM].ref_off + c] = wb[c];
        w[s.seg[SEG_BN_RV].ref_off + c] = wb[s.d3 + c];
      }
      w[s.seg[SEG_BN_NBT].ref_off] = wb[2 * s.d3];
    }
  }
  if (sc) {   // one staged struct, one checked copy
    CU(launch_scalars(h->dS, chain, h->stage_sc, 0, h->stream));
    h->launches++;
    CU(cudaMemcpyAsync(sc, h->stage_sc, sizeof(bae_chain_scalars), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return BAE_OK;
}

int bae_set_state(bae_handle* h, int chain, const float* theta, const float* w, const float* m, const float* v,
                  const bae_chain_scalars* sc) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  int rc;
  const size_t o = (size_t)chain * s.Pp;
  if (theta && (rc = put_vec(h, theta, s.theta + o))) return rc;
  if (m && (rc = put_vec(h, m, s.m + o))) return rc;
  if (v && (rc = put_vec(h, v, s.v + o))) return rc;
  if (w) {
    std::vector<float> wb(s.nbuf);
    for (int c = 0; c < s.d3; ++c) {
      wb[c] = w[s.seg[SEG_BN_RM].ref_off + c];
      wb[s.d3 + c] = w[s.seg[SEG_BN_RV].ref_off + c];
    }
    wb[2 * s.d3] = w[s.seg[SEG_BN_NBT].ref_off];
    CU(cudaMemcpyAsync(s.wb + (size_t)chain * s.nbuf, wb.data(), sizeof(float) * s.nbuf, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  if (sc) {
    h->host_step = sc->step;   // all chains of a handle advance together (bae_step / bae_run)
    CU(cudaMemcpyAsync(h->stage_sc, sc, sizeof(bae_chain_scalars), cudaMemcpyHostToDevice, h->stream));
    CU(launch_scalars(h->dS, chain, h->stage_sc, 1, h->stream));
    h->launches++;
    CU(cudaStreamSynchronize(h->stream));   // `sc` is caller-owned pageable memory
  }
  return BAE_OK;
}

// Issues the per-phase launches of `n_rep` iterations of phases [p_begin, p_end): the step phases (prologue .. epilogue)
// chain group after chain group (the activation buffers hold G chains), the exchange phases once over all chains.
static cudaError_t emit_phases(bae_handle* h, int p_begin, int p_end, int n_rep, int force_swap, int* n_launch) {
  const int C = h->hs.C, G = h->hs.G;
  const int ps_end = std::min(p_end, (int)P_SWAP);
  cudaError_t e = cudaSuccess;
  const int64_t n0 = h->n_emit;
  // 3xFP16 split: whatever rewrote theta since the last program (initialisation, set_state, an exchange) - recompute the
  // weight bounds of every chain once in front of the program
  auto mark = [&](int p) {
    if (!h->phase_times) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, h->stream);
    h->pt_events.emplace_back(p, ev);
  };
  mark(-1);
  if (h->hs.f16 && p_begin == P_PRO) { e = launch_one(h, P_WAMAX0, 0, C, 1); mark(P_WAMAX0); }
  for (int r = 0; r < n_rep && e == cudaSuccess; ++r) {
    if (p_begin < ps_end)
      for (int cb = 0; cb < C && e == cudaSuccess; cb += G)
        for (int p = p_begin; p < ps_end && e == cudaSuccess; ++p) {
          if (p == P_WAMAX0 && r == 0 && p_begin == P_PRO) continue;   // the forced launch above has just done this
          e = launch_one(h, p, cb, std::min(cb + G, C), force_swap);
          mark(p);
        }
    for (int p = std::max(p_begin, (int)P_SWAP); p < p_end && e == cudaSuccess; ++p) { e = launch_one(h, p, 0, C, force_swap); mark(p); }
  }
  if (n_launch) *n_launch = (int)(h->n_emit - n0);
  return e;
}

static int coop_launch(bae_handle* h, int p_begin, int p_end, int n_rep, int force_swap) {
  if (!h->data_loaded) return fail(BAE_ERR_STATE, "upload data first");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->ev0, h->stream));
  const int C = h->hs.C, G = h->hs.G;
  if (h->narrow && p_begin == P_PRO && p_end > P_EPI) {
    // narrow-width shapes: the iterations between two exchange rounds are ONE launch of the cluster kernel; the exchange
    // sweep (gated on the device's step counter like everywhere) is the generic swap phase program
    const bool with_swap = p_end > P_SWAP;
    const int si = h->hs.swap_interval;
    int left = n_rep;
    while (left > 0) {
      const int k = with_swap ? (int)std::min<long long>(left, si - h->host_step % si) : left;
      CU(launch_narrow(h->dS, C, h->narrow_cluster, k, h->p_used, h->stream));
      h->launches++;
      h->host_step += k;
      left -= k;
      if (with_swap && h->host_step % si == 0) {
        if (C <= 16) CU(launch_swap_cta(h->dS, force_swap, h->stream));   // the three exchange phases in one CTA
        else CU(launch_program(h->dS, 0, h->grid, P_SWAP, P_END, 1, 0, C, force_swap, h->stream));
        h->launches++;
      }
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    return BAE_OK;
  }
  if (p_begin == P_PRO && p_end > P_EPI) h->host_step += n_rep;
  if (h->phase_launch) {
    // One launch per phase (the exchange phases test the step counter on the device, like the persistent kernel); the
    // launches of a call are captured once into a CUDA graph and replayed.
    const int groups = (C + G - 1) / G;
    const int n_est = n_rep * (p_end - p_begin) * groups;
    int n_launch = 0;
    const std::array<int, 4> key = {p_begin, p_end, n_rep, force_swap};
    // at most 32 distinct call shapes are kept as graphs (a caller that varies n_steps freely falls back to direct launches)
    if (h->use_graph && n_est <= 8192 && (h->graphs.count(key) || h->graphs.size() < 32)) {
      auto it = h->graphs.find(key);
      if (it == h->graphs.end()) {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const cudaError_t e = emit_phases(h, p_begin, p_end, n_rep, force_swap, &n_launch);
        const cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
        if (e != cudaSuccess || e2 != cudaSuccess) {
          if (graph) cudaGraphDestroy(graph);
          return fail(BAE_ERR_CUDA, std::string("graph capture of the step program failed: ") +
                                        cudaGetErrorString(e != cudaSuccess ? e : e2));
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e3 = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e3 != cudaSuccess) return fail(BAE_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e3));
        it = h->graphs.emplace(key, exec).first;
        h->graph_launches[key] = n_launch;
      }
      n_launch = h->graph_launches[key];
      CU(cudaEventRecord(h->ev0, h->stream));
      CU(cudaGraphLaunch(it->second, h->stream));
    } else {
      CU(emit_phases(h, p_begin, p_end, n_rep, force_swap, &n_launch));
    }
    h->launches += n_launch;
  } else if (G >= C) {
    CU(launch_program(h->dS, h->hs.tc, h->grid, p_begin, p_end, n_rep, 0, C, force_swap, h->stream));
    h->launches++;
  } else {
    // persistent kernel with chain groups: the step phases of an iteration group after group, then the exchange phases
    const int ps_end = std::min(p_end, (int)P_SWAP);
    const bool swap_inside = p_end > P_SWAP;
    for (int r = 0; r < (swap_inside ? n_rep : 1); ++r) {
      if (p_begin < ps_end)
        for (int cb = 0; cb < C; cb += G) {
          CU(launch_program(h->dS, h->hs.tc, h->grid, p_begin, ps_end, swap_inside ? 1 : n_rep, cb, std::min(cb + G, C), force_swap, h->stream));
          h->launches++;
        }
      if (swap_inside) {
        CU(launch_program(h->dS, h->hs.tc, h->grid, std::max(p_begin, (int)P_SWAP), p_end, 1, 0, C, force_swap, h->stream));
        h->launches++;
      }
    }
  }
  CU(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return BAE_OK;
}

int bae_step(bae_handle* h, int n_steps) {
  if (!h || n_steps < 0) return fail(BAE_ERR_INVALID, "bad argument");
  if (n_steps == 0) return BAE_OK;
  return coop_launch(h, P_PRO, P_EPI + 1, n_steps, 0);
}

int bae_swap_local(bae_handle* h) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  return coop_launch(h, P_SWAP, P_END, 1, 1);
}

int bae_run(bae_handle* h, int n_steps) {
  if (!h || n_steps < 0) return fail(BAE_ERR_INVALID, "bad argument");
  if (n_steps == 0) return BAE_OK;
  if (h->hs.shard_world > 1)
    return fail(BAE_ERR_STATE, "this handle holds a block of a sharded ladder: use bae_step + bae_swap_exchange");
  return coop_launch(h, P_PRO, P_END, n_steps, 0);
}

int bae_last_step_ms(bae_handle* h, float* ms) {
  if (!h || !ms) return fail(BAE_ERR_INVALID, "null argument");
  if (!h->timed) return fail(BAE_ERR_STATE, "no launch recorded");
  CU(cudaEventSynchronize(h->ev1));
  CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return BAE_OK;
}

int bae_eval(bae_handle* h, int chain, int which, const float* w, double* sse, double* mse, float* grads, float* out) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  if (!h->data_loaded) return fail(BAE_ERR_STATE, "upload data first");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  if (which != 0 && which != 1) return fail(BAE_ERR_INVALID, "which must be 0 (train) or 1 (test)");
  if (grads && which != 0) return fail(BAE_ERR_INVALID, "gradients are defined on the training matrix only (MAIN:473)");
  CU(cudaSetDevice(h->device));
  const int E = s.C;  // scratch chain slot
  float* th = s.theta + (size_t)E * s.Pp;
  int rc;
  if (w) {
    if ((rc = put_vec(h, w, th))) return rc;
  } else {
    CU(cudaMemcpyAsync(th, s.theta + (size_t)chain * s.Pp, sizeof(float) * s.Pp, cudaMemcpyDeviceToDevice, h->stream));
  }
  const int pass = which == 0 ? 0 : 2;
  const int rows = which == 0 ? s.n_train : s.n_test;
  double a = 0;
  if ((rc = forward_one(h, E, pass, &a, nullptr))) return rc;
  if (sse) *sse = a;
  if (mse) *mse = a / ((double)rows * s.d0);
  if (out) {
    // the loss epilogue stores (out - X) * scale; undo it on the host
    std::vector<float> d((size_t)rows * s.ld0), x((size_t)rows * s.ld0);
    CU(cudaMemcpyAsync(d.data(), s.DOUT, sizeof(float) * d.size(), cudaMemcpyDeviceToHost, h->stream));   // the probe ran alone: slot 0
    CU(cudaMemcpyAsync(x.data(), which == 0 ? s.Xtrain : s.Xtest, sizeof(float) * x.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));

    const double scale = 2.0 / ((double)rows * s.d0);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < s.d0; ++c)
        out[(size_t)r * s.d0 + c] = (float)((double)d[(size_t)r * s.ld0 + c] / scale + (double)x[(size_t)r * s.ld0 + c]);
  }
  if (grads) {
    for (int k = 0; k < s.max_ksplit; ++k)
      CU(cudaMemsetAsync(s.grad + ((size_t)k * s.C_alloc + E) * s.Pp, 0, sizeof(float) * s.Pp, h->stream));
    if ((rc = run_phases_plain(h, P_BWD1, P_BWD1 + 7, E, E + 1))) return rc;
    if (s.max_ksplit > 1) CU(launch_grad_reduce(h->dS, E, h->stream));
    if ((rc = get_vec(h, s.grad + (size_t)E * s.Pp, grads))) return rc;
  }
  return BAE_OK;
}

int bae_eval_bn_stats(bae_handle* h, int which, float* mean, float* var) {
  if (!h || !mean || !var) return fail(BAE_ERR_INVALID, "null argument");
  if (which != 0 && which != 1) return fail(BAE_ERR_INVALID, "which must be 0 (train) or 1 (test)");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  const size_t slot = which == 0 ? 0 : 2, E = s.C;   // the probe's scratch chain
  std::vector<float> mu(s.ld3), va(s.ld3);
  CU(cudaMemcpyAsync(mu.data(), s.bn_mu + (slot * s.C_alloc + E) * s.ld3, sizeof(float) * s.ld3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(va.data(), s.bn_var + (slot * s.C_alloc + E) * s.ld3, sizeof(float) * s.ld3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < s.d3; ++c) { mean[c] = mu[c]; var[c] = va[c]; }
  return BAE_OK;
}

int bae_forward_all(bae_handle* h, int which, float* features, float* recon, double* sse) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  if (!h->data_loaded) return fail(BAE_ERR_STATE, "upload data first");
  if (which != 0 && which != 1) return fail(BAE_ERR_INVALID, "which must be 0 (train) or 1 (test)");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  // the second-pass train forward phases / the test forward phases run for every chain (needLang = 0); they write
  // activations, batch statistics and loss partials only
  const int p0 = which == 0 ? P_FWD2 : P_FWDT, pass = which == 0 ? 1 : 2;
  const int rows = which == 0 ? s.n_train : s.n_test;
  const float* X = which == 0 ? s.Xtrain : s.Xtest;
  if (!h->stage_rows && (features || recon)) CU(dalloc(h, &h->stage_rows, (size_t)s.n_max * std::max(s.d0, s.d3)));
  const float inv_scale = (float)(((double)rows * s.d0) / 2.0);
  for (int cb = 0; cb < s.C; cb += s.G) {       // chain group after chain group (the activation buffers hold G chains)
    const int ce = std::min(cb + s.G, s.C);
    int rc2 = run_phases_plain(h, P_WAMAX0, P_WAMAX0 + 1, cb, ce);
    if (rc2) return rc2;
    if ((rc2 = run_phases_plain(h, p0, p0 + 7, cb, ce))) return rc2;
    for (int c = cb; c < ce; ++c) {
      const size_t slot = (size_t)(c - cb);
      if (features) {
        CU(launch_extract(s.Z + slot * s.n_max * s.ld3, s.ld3, nullptr, 0, rows, s.d3, 0.f, 0, h->stage_rows, h->stream));
        h->launches++;
        CU(cudaMemcpyAsync(features + (size_t)c * rows * s.d3, h->stage_rows, sizeof(float) * rows * s.d3, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));   // pageable destination; the staging buffer is reused
      }
      if (recon) {
        CU(launch_extract(s.DOUT + slot * s.n_max * s.ld0, s.ld0, X, s.ld0, rows, s.d0, inv_scale, 1, h->stage_rows, h->stream));
        h->launches++;
        CU(cudaMemcpyAsync(recon + (size_t)c * rows * s.d0, h->stage_rows, sizeof(float) * rows * s.d0, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
      }
    }
  }
  if (sse) {
    std::vector<double> part((size_t)s.C * s.loss_stride * 2);
    CU(cudaMemcpyAsync(part.data(), s.loss_part + (size_t)pass * s.C_alloc * s.loss_stride * 2, sizeof(double) * part.size(),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < s.C; ++c) {
      double a = 0;
      for (int i = 0; i < s.loss_stride; ++i) a += part[((size_t)c * s.loss_stride + i) * 2];
      sse[c] = a;
    }
  }
  return BAE_OK;
}

int bae_debug_draws(bae_handle* h, int chain, int step, float* noise, double* scalars3) {
  if (!h || !noise || !scalars3) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  CU(launch_debug_draws(h->dS, chain, step, h->stage_noise, h->stage_d, h->stream));
  h->launches++;
  CU(cudaMemcpyAsync(noise, h->stage_noise, sizeof(float) * s.w_size, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(scalars3, h->stage_d, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

int bae_debug_swap_uniform(bae_handle* h, int ensemble, int round, int pair, double* u) {
  if (!h || !u) return fail(BAE_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  CU(launch_debug_swap_u(h->dS, h->hs.chain_id_base / h->hs.n_rungs + ensemble, round, pair, h->stage_d, h->stream));
  h->launches++;
  CU(cudaMemcpyAsync(u, h->stage_d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

int bae_read_traces(bae_handle* h, int chain, int first_step, int n, bae_trace* outp) {
  if (!h || !outp) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C || first_step < 0 || n < 0 || first_step + n > s.samples)
    return fail(BAE_ERR_INVALID, "trace range out of bounds");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(outp, s.traces + (size_t)chain * s.samples + first_step, sizeof(bae_trace) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

// The same records of EVERY resident chain in one strided copy and one synchronisation: out[chain][n]
int bae_read_traces_all(bae_handle* h, int first_step, int n, bae_trace* outp) {
  if (!h || !outp) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  if (first_step < 0 || n < 0 || first_step + n > s.samples) return fail(BAE_ERR_INVALID, "trace range out of bounds");
  if (n == 0) return BAE_OK;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpy2DAsync(outp, sizeof(bae_trace) * n, s.traces + first_step, sizeof(bae_trace) * s.samples, sizeof(bae_trace) * n, s.C,
                       cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

int bae_swap_counts(bae_handle* h, int64_t* num_swap, int64_t* total) {
  if (!h || !num_swap || !total) return fail(BAE_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  long long a = 0, b = 0;
  CU(cudaMemcpyAsync(&a, h->hs.num_swap, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(&b, h->hs.total_swap, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  *num_swap = a;
  *total = b;
  return BAE_OK;
}

// ---- ladder sharded over ranks ------------------------------------------------------------------------------------
namespace {
struct Nccl {
  int (*all_gather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*group_start)() = nullptr;
  int (*group_end)() = nullptr;
  const char* (*error_string)(int) = nullptr;
  int (*comm_count)(void*, int*) = nullptr;
  int (*comm_user_rank)(void*, int*) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat = 7, kNcclDouble = 8;   // ncclFloat32 / ncclFloat64 of nccl.h

Nccl* nccl_api() {
  static Nccl api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  // the caller created its communicator with the libnccl.so.2 already loaded into this process: bind to that one
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return &api;
  api.all_gather = reinterpret_cast<decltype(api.all_gather)>(dlsym(lib, "ncclAllGather"));
  api.send = reinterpret_cast<decltype(api.send)>(dlsym(lib, "ncclSend"));
  api.recv = reinterpret_cast<decltype(api.recv)>(dlsym(lib, "ncclRecv"));
  api.group_start = reinterpret_cast<decltype(api.group_start)>(dlsym(lib, "ncclGroupStart"));
  api.group_end = reinterpret_cast<decltype(api.group_end)>(dlsym(lib, "ncclGroupEnd"));
  api.error_string = reinterpret_cast<decltype(api.error_string)>(dlsym(lib, "ncclGetErrorString"));
  api.comm_count = reinterpret_cast<decltype(api.comm_count)>(dlsym(lib, "ncclCommCount"));
  api.comm_user_rank = reinterpret_cast<decltype(api.comm_user_rank)>(dlsym(lib, "ncclCommUserRank"));
  api.ok = api.all_gather && api.send && api.recv && api.group_start && api.group_end && api.comm_count && api.comm_user_rank;
  return &api;
}
}  // namespace
#define NC(x)                                                                                                     \
  do {                                                                                                            \
    int r_ = (x);                                                                                                 \
    if (r_ != 0) return fail(BAE_ERR_CUDA, std::string(#x) + ": NCCL error " + (nc->error_string ? nc->error_string(r_) : "?")); \
  } while (0)

int bae_shard_config(bae_handle* h, int rank, int world) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  if (world < 1 || rank < 0 || rank >= world) return fail(BAE_ERR_INVALID, "bad rank / world");
  DevState& s = h->hs;
  const int L = s.n_rungs, R = L * world, E = s.C / L;
  if (s.chain_id_base % R != 0) return fail(BAE_ERR_INVALID, "chain_id_base must be a multiple of the ladder length n_rungs * world");
  CU(cudaSetDevice(h->device));
  s.shard_rank = rank; s.shard_world = world;
  std::vector<int> gid(s.C_alloc);
  for (int e = 0; e < E; ++e)
    for (int l = 0; l < L; ++l) gid[e * L + l] = s.chain_id_base + e * R + rank * L + l;   // the ladder position keys the streams
  gid[s.C] = s.chain_id_base + E * R + rank;
  CU(cudaMemcpyAsync(s.ch.gid, gid.data(), sizeof(int) * s.C_alloc, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (!s.xt_local) {
    CU(dalloc(h, &s.xt_local, (size_t)s.C * 5));
    CU(dalloc(h, &s.xt_all, (size_t)world * s.C * 5));
    CU(dalloc(h, &s.xsrc, (size_t)E * R));
    CU(cudaMallocHost(reinterpret_cast<void**>(&h->xsrc_host), sizeof(int32_t) * E * R));
  }
  int rc = upload_state(h);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

int bae_plan_exchange(const int32_t* src, int E, int R, int world, int rank, int32_t* out, int max_ops) {
  if (!src || !out || E < 1 || R < 1 || world < 1 || R % world != 0 || rank < 0 || rank >= world)
    return fail(BAE_ERR_INVALID, "bad argument");
  const int L = R / world;
  int n = 0;
  // both sides walk (ensemble, destination rung) in ascending order: the k-th message from rank a to rank b is the same
  // configuration on both, which is all NCCL's point-to-point matching needs
  for (int e = 0; e < E; ++e)
    for (int p = 0; p < R; ++p) {
      const int sr = src[(size_t)e * R + p];
      if (sr < 0 || sr >= R) return fail(BAE_ERR_INVALID, "permutation entry out of range");
      const int po = p / L, so = sr / L;
      if (po == so) continue;
      if (so != rank && po != rank) continue;
      if (n >= max_ops) return fail(BAE_ERR_INVALID, "operation buffer too small");
      out[4 * n + 0] = so == rank ? 0 : 1;
      out[4 * n + 1] = so == rank ? po : so;
      out[4 * n + 2] = e;
      out[4 * n + 3] = so == rank ? sr : p;
      ++n;
    }
  return n;
}

int64_t bae_exchange_vectors_sent(const bae_handle* h) { return h ? h->vectors_sent : 0; }

int bae_swap_exchange(bae_handle* h, void* comm) {
  if (!h || !comm) return fail(BAE_ERR_INVALID, "null argument");
  DevState& s = h->hs;
  if (s.shard_world < 2 || !s.xt_local) return fail(BAE_ERR_STATE, "call bae_shard_config(rank, world >= 2) first");
  Nccl* nc = nccl_api();
  if (!nc->ok) return fail(BAE_ERR_CUDA, "libnccl.so.2 is not loadable in this process");
  int cnt = 0, urank = -1;
  NC(nc->comm_count(comm, &cnt));
  NC(nc->comm_user_rank(comm, &urank));
  if (cnt != s.shard_world || urank != s.shard_rank) return fail(BAE_ERR_INVALID, "communicator does not match bae_shard_config(rank, world)");
  CU(cudaSetDevice(h->device));
  const int L = s.n_rungs, R = L * s.shard_world, E = s.C / L, rank = s.shard_rank;
  cudaStream_t st = h->stream;
  CU(launch_xtable(h->dS, st));
  NC(nc->all_gather(s.xt_local, s.xt_all, (size_t)s.C * 5, kNcclDouble, comm, st));
  CU(launch_xsweep(h->dS, st));
  CU(cudaMemcpyAsync(h->xsrc_host, s.xsrc, sizeof(int32_t) * E * R, cudaMemcpyDeviceToHost, st));
  CU(launch_xgather(h->dS, 2 * h->n_sms, st));     // same-GPU moves overlap the host's planning below
  h->launches += 3;
  CU(cudaStreamSynchronize(st));                   // the one host synchronisation of the round: the permutation
  h->xops.resize((size_t)4 * 2 * E * R);
  const int nops = bae_plan_exchange(h->xsrc_host, E, R, s.shard_world, rank, h->xops.data(), 2 * E * R);
  if (nops < 0) return nops;
  if (nops > 0) {
    NC(nc->group_start());
    for (int k = 0; k < nops; ++k) {
      const int kind = h->xops[4 * k], peer = h->xops[4 * k + 1], e = h->xops[4 * k + 2], rung = h->xops[4 * k + 3];
      const size_t c = (size_t)e * L + (rung - rank * L);
      if (kind == 0) {
        NC(nc->send(s.theta + c * s.Pp, (size_t)s.Pp, kNcclFloat, peer, comm, st));
        NC(nc->send(s.wb + c * s.nbuf, (size_t)s.nbuf, kNcclFloat, peer, comm, st));
        h->vectors_sent++;
      } else {
        NC(nc->recv(s.wcur + c * s.Pp, (size_t)s.Pp, kNcclFloat, peer, comm, st));
        NC(nc->recv(s.wb_tmp + c * s.nbuf, (size_t)s.nbuf, kNcclFloat, peer, comm, st));
      }
    }
    NC(nc->group_end());
  }
  CU(launch_xscatter(h->dS, 2 * h->n_sms, st));
  h->launches++;
  return BAE_OK;
}

int bae_last_sweep(bae_handle* h, int32_t* src_out) {
  if (!h || !src_out) return fail(BAE_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(src_out, h->hs.swap_src, sizeof(int32_t) * h->hs.C, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}

// (eta, stored likelihood, adapttemp, SSE of `w`) of every resident chain: the rows a rank contributes to the
// all-gather of an exchange round (MAIN:594-597 posts these scalars next to the weights). out: [n_chains][4].
int bae_exchange_table(bae_handle* h, double* out) {
  if (!h || !out) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  std::vector<double> col((size_t)4 * s.C);
  const double* src[4] = {s.ch.eta, s.ch.lik, s.ch.adapt, s.ch.last_sse};
  for (int k = 0; k < 4; ++k)
    CU(cudaMemcpyAsync(col.data() + (size_t)k * s.C, src[k], sizeof(double) * s.C, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < s.C; ++c)
    for (int k = 0; k < 4; ++k) out[(size_t)c * 4 + k] = col[(size_t)k * s.C + c];
  return BAE_OK;
}

int bae_exchange_pack(bae_handle* h, int chain, float* dev_w, double* host_scalars4) {
  if (!h || !dev_w) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  CU(launch_unpack(h->dS, s.theta + (size_t)chain * s.Pp, dev_w, h->stream));
  h->launches++;
  int alias = 1;
  int rc = get1<int>(h, s.ch.alias, chain, &alias);
  if (rc) return rc;
  if (!alias) {  // detached `w`: its BatchNorm buffers live in wb
    CU(cudaMemcpyAsync(dev_w + s.seg[SEG_BN_RM].ref_off, s.wb + (size_t)chain * s.nbuf, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(dev_w + s.seg[SEG_BN_RV].ref_off, s.wb + (size_t)chain * s.nbuf + s.d3, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(dev_w + s.seg[SEG_BN_NBT].ref_off, s.wb + (size_t)chain * s.nbuf + 2 * s.d3, sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (host_scalars4) {   // optional: bae_exchange_table reads them for all chains at once
    get1(h, s.ch.eta, chain, &host_scalars4[0]);
    get1(h, s.ch.lik, chain, &host_scalars4[1]);
    get1(h, s.ch.adapt, chain, &host_scalars4[2]);
    get1(h, s.ch.last_sse, chain, &host_scalars4[3]);
  }
  return BAE_OK;
}

int bae_exchange_unpack(bae_handle* h, int chain, const float* dev_w, const double* host_scalars2) {
  if (!h || !dev_w || !host_scalars2) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  if (chain < 0 || chain >= s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  // buffers of the received dict go to wb; trainables go straight into the live model (reloaded next step anyway)
  CU(cudaMemcpyAsync(s.wb + (size_t)chain * s.nbuf, dev_w + s.seg[SEG_BN_RM].ref_off, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(s.wb + (size_t)chain * s.nbuf + s.d3, dev_w + s.seg[SEG_BN_RV].ref_off, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(s.wb + (size_t)chain * s.nbuf + 2 * s.d3, dev_w + s.seg[SEG_BN_NBT].ref_off, sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  // keep the live buffers: pack into scratch, then copy trainable segments only
  float* scratch = s.wcur + (size_t)chain * s.Pp;
  CU(cudaMemsetAsync(scratch, 0, sizeof(float) * s.Pp, h->stream));
  CU(launch_pack(h->dS, dev_w, scratch, h->stream));
  h->launches++;
  for (int i = 0; i < kNumSegs; ++i) {
    const Seg& g = s.seg[i];
    if (!g.trainable) continue;
    CU(cudaMemcpyAsync(s.theta + (size_t)chain * s.Pp + g.pad_off, scratch + g.pad_off, sizeof(float) * g.rows * g.ld,
                       cudaMemcpyDeviceToDevice, h->stream));
  }

  put1(h, s.ch.eta, chain, host_scalars2[0]);
  put1(h, s.ch.last_sse, chain, host_scalars2[1]);
  put1<int>(h, s.ch.alias, chain, 0);
  return BAE_OK;
}

// Host-side restatement of run_swap_decide for ONE ensemble (MAIN:962-1011, 1059-1065, with the message-slot quirk of
// MAIN:999-1005): the sweep every rank of a ladder-sharded run executes on the all-gathered table. No GPU involved.
// table: [n_rungs][4] = (eta, stored likelihood, adapttemp, SSE of `w`) by ladder position; uniforms: [n_rungs - 1].
// src_out[p] = position whose configuration ends at position p.
int bae_host_sweep(const double* table, int n_rungs, double n_elems, const float* uniforms, int* src_out, int* n_swapped) {
  if (!table || !uniforms || !src_out || n_rungs < 1) return fail(BAE_ERR_INVALID, "bad argument");
  auto loglik = [&](double sse, double tau_sq, double temp) {
    return (-(n_elems * 0.5) * std::log(2.0 * 3.14159265358979323846 * tau_sq) - 0.5 * tau_sq * sse) / temp;
  };
  int cfg_prev = 0, swapped = 0;
  double lh_prev = table[1], T_prev = table[2];
  for (int p = 0; p + 1 < n_rungs; ++p) {
    const int cfg1 = cfg_prev, cfg2 = p + 1;
    const double lhood1 = lh_prev, T1 = T_prev;
    const double lhood2 = table[cfg2 * 4 + 1], T2 = table[cfg2 * 4 + 2];
    const double lhood12 = loglik(table[cfg1 * 4 + 3], std::exp(table[cfg1 * 4 + 0]), T2);
    const double lhood21 = loglik(table[cfg2 * 4 + 3], std::exp(table[cfg2 * 4 + 0]), T1);
    const double ex = (lhood12 - lhood1) + (lhood21 - lhood2);
    const double prob = std::fmin(1.0, std::exp(ex));   // MAIN:988 min(1, exp(.)); a NaN exponent gives 1, as there and on the device
    if ((double)uniforms[p] < prob) {
      ++swapped;
      src_out[p] = cfg2;
      cfg_prev = cfg1; lh_prev = lhood12; T_prev = T1;
    } else {
      src_out[p] = cfg1;
      cfg_prev = cfg2; lh_prev = lhood2; T_prev = T2;
    }
  }
  src_out[n_rungs - 1] = cfg_prev;
  if (n_swapped) *n_swapped = swapped;
  return BAE_OK;
}

// All moves of an exchange round that stay on this GPU, in one gather + one scatter pass of the device's own swap
// phases: position p receives the configuration of local chain src[p] (src[p] == p: keeps its own, or will be
// overwritten by bae_exchange_unpack with a configuration from another rank). Like bae_exchange_finish, every
// chain's `w` ends as a detached dict (MAIN:602).
int bae_exchange_local(bae_handle* h, const int* src) {
  if (!h || !src) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  for (int c = 0; c < s.C; ++c)
    if (src[c] < 0 || src[c] >= s.C) return fail(BAE_ERR_INVALID, "source chain out of range");
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(s.swap_src, src, sizeof(int) * s.C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));   // `src` is caller-owned pageable memory
  return run_phases_plain(h, P_SWAP + 1, P_SWAP + 3, 0, s.C);
}

int bae_exchange_finish(bae_handle* h) {
  if (!h) return fail(BAE_ERR_INVALID, "null handle");
  const DevState& s = h->hs;
  CU(cudaSetDevice(h->device));
  // every chain's `w` becomes a detached dict (MAIN:602); chains that kept their own configuration
  // snapshot their live buffers
  std::vector<int> alias(s.C);
  CU(cudaMemcpyAsync(alias.data(), s.ch.alias, sizeof(int) * s.C, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < s.C; ++c) {
    if (!alias[c]) continue;
    const float* th = s.theta + (size_t)c * s.Pp;
    CU(cudaMemcpyAsync(s.wb + (size_t)c * s.nbuf, th + s.seg[SEG_BN_RM].pad_off, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(s.wb + (size_t)c * s.nbuf + s.d3, th + s.seg[SEG_BN_RV].pad_off, sizeof(float) * s.d3, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(s.wb + (size_t)c * s.nbuf + 2 * s.d3, th + s.seg[SEG_BN_NBT].pad_off, sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    alias[c] = 0;
  }
  std::vector<int> zero(s.C, 0);
  CU(cudaMemcpyAsync(s.ch.alias, zero.data(), sizeof(int) * s.C, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return BAE_OK;
}


// Debug probe: copy a per-chain device array (padded layout, hi part on the tcgen05 path) to the host.
// `chain` may be n_chains (the parity-probe scratch slot). Returns the row pitch in *ld.
int bae_debug_buffer(bae_handle* h, const char* name, int chain, float* out, int64_t n_floats, int* ld, int lo_part) {
  if (!h || !name || !out) return fail(BAE_ERR_INVALID, "null argument");
  const DevState& s = h->hs;
  if (chain < 0 || chain > s.C) return fail(BAE_ERR_INVALID, "chain out of range");
  CU(cudaSetDevice(h->device));
  struct E { const char* n; const float* hi; const float* lo; int ld; };
  const E tab[] = {{"H1", s.H1, nullptr, s.ld1}, {"H2", s.H2, nullptr, s.ld2}, {"Z", s.Z, nullptr, s.ld3}, {"Y", s.Y, nullptr, s.ld3},
                   {"H4", s.H4, nullptr, s.ld2}, {"H5", s.H5, nullptr, s.ld1}, {"DOUT", s.DOUT, nullptr, s.ld0},
                   {"D5", s.D5, nullptr, s.ld1}, {"D4", s.D4, nullptr, s.ld2}, {"DY", s.DY, nullptr, s.ld3},
                   {"D2", s.D2, nullptr, s.ld2}, {"D1", s.D1, nullptr, s.ld1}};
  for (const E& e : tab) {
    if (strcmp(e.n, name) != 0) continue;
    const float* src = lo_part ? e.lo : e.hi;
    if (!src) return fail(BAE_ERR_INVALID, "no such part");
    const size_t per = (size_t)s.n_max * e.ld;
    if ((size_t)n_floats > per) return fail(BAE_ERR_INVALID, "too many floats");
    // activation buffers hold the last chain group that ran; the parity probe's scratch chain (chain == n_chains) ran alone
    const size_t slot = chain == s.C ? 0 : (size_t)(chain % s.G);
    CU(cudaMemcpyAsync(out, src + slot * per, sizeof(float) * n_floats, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (ld) *ld = e.ld;
    return BAE_OK;
  }
  if (strcmp(name, "grad") == 0 || strcmp(name, "theta") == 0) {
    const float* src = name[0] == 'g' ? s.grad : s.theta;
    if (!src) return fail(BAE_ERR_INVALID, "no such array");
    if (n_floats > s.Pp) return fail(BAE_ERR_INVALID, "too many floats");
    CU(cudaMemcpyAsync(out, src + (size_t)chain * s.Pp, sizeof(float) * n_floats, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (ld) *ld = 0;
    return BAE_OK;
  }
  return fail(BAE_ERR_INVALID, "unknown buffer name");
}


// Debug: timeline trace of CTA 0 (needs BAE_TC_TRACE=1 at create time): 4096 int64 values.
int bae_debug_trace(bae_handle* h, long long* out, int clear) {
  if (!h || !out) return fail(BAE_ERR_INVALID, "null argument");
  if (!h->hs.dbg) return fail(BAE_ERR_STATE, "tracing not enabled (BAE_TC_TRACE=1)");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(out, h->hs.dbg, 4096 * sizeof(long long), cudaMemcpyDeviceToHost));
  if (clear) CU(cudaMemset(h->hs.dbg, 0, 4096 * sizeof(long long)));
  return BAE_OK;
}

// Debug (BAE_PHASE_TIMES=1): milliseconds spent per phase index (the amax-zeroing / image launches that belong to a phase
// included) since the last call, measured with an event after every phase of back-to-back direct launches; ms64 / n64: [64].
int bae_debug_phase_times(bae_handle* h, double* ms64, int* n64) {
  if (!h || !ms64 || !n64) return fail(BAE_ERR_INVALID, "null argument");
  CU(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 64; ++i) { ms64[i] = 0.0; n64[i] = 0; }
  for (size_t i = 1; i < h->pt_events.size(); ++i) {
    const int p = h->pt_events[i].first;
    if (p < 0 || p >= 64) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->pt_events[i - 1].second, h->pt_events[i].second) == cudaSuccess) { ms64[p] += ms; n64[p]++; }
  }
  for (auto& pe : h->pt_events) cudaEventDestroy(pe.second);
  h->pt_events.clear();
  return BAE_OK;
}

// Debug: cudaProfilerStart / cudaProfilerStop around a region (ncu --profile-from-start off).
int bae_debug_profiler(int on) {
  CU(on ? cudaProfilerStart() : cudaProfilerStop());
  return BAE_OK;
}

}  // extern "C"
