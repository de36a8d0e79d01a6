// tcgen05 / TMEM / TMA tile engine (sm_100a): chain-batched GEMM tiles with 3xTF32 split operands.
//
//   D[128 x BN] (FP32, TMEM) += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        per 8-wide K step
//
// Operands live in HBM as plain FP32. The tensor core reads a 32-bit container and ignores the low 13
// mantissa bits (measured: tests/cuda/tc_selftest.cu), so the raw tile IS the `hi` operand; two converter
// warps compute lo = x - trunc_tf32(x) (exact in FP32) into a second shared-memory tile. Three kind::tf32
// MMAs per K step then recover ~FP32 products with FP32 accumulation in TMEM. Raw tiles (16 floats of K per
// stage) are staged by TMA into a 4-stage shared-memory ring; one elected thread issues tcgen05.mma; eight
// epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue (bias / sigmoid / loss /
// sigmoid-derivative / gradient routing) with coalesced, shared-memory-transposed loads and stores.
//
// Warp roles inside the 384-thread CTA: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-3 = hi/lo
// converters, warps 4-11 = epilogue (two warps per TMEM lane quarter, alternating 32-column chunks).
#pragma once
#include <cuda.h>

#include "bae_common.cuh"

namespace bae {
namespace tc {

constexpr int kThreadsTc = 384;
constexpr int kStages = 5;                     // raw-tile ring fed by TMA
constexpr int kLoStages = 3;                   // lo-tile ring written by the converter warps
constexpr int kBM = 128;
constexpr int kBK = 16;                        // floats of K per stage (64-byte rows: SWIZZLE_64B for K-major tiles)
constexpr int kMaxBN = 240;
constexpr int kATileBytes = kBM * kBK * 4;     // 8 KB
constexpr int kBTileBytes = kMaxBN * kBK * 4;  // 15 KB
constexpr int kMnChunkBytes = kBK * 128;       // MN-major chunk: [16 K-rows][32 floats]
constexpr int kStageBytes = kATileBytes + kBTileBytes;          // one raw stage (A, B) or one lo stage = 23 KB
constexpr int kEpiPitch = 36;                  // floats; 144-byte rows keep float4 accesses conflict-free
constexpr int kEpiWarps = 8;
constexpr int kConvWarps = 10;                 // warps 2..11 convert; warps 4..11 then run the epilogue of the tile
constexpr int kEpiBytes = kEpiWarps * 32 * kEpiPitch * 4;
constexpr int kSmemBytes = (kStages + kLoStages) * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/ + kEpiBytes;
// stage layout: [A 8 KB][B 15 KB]; kStages raw stages followed by kLoStages lo stages
constexpr int kTmemCols = 512;                 // (tcMain + 1) accumulators of BN columns each, single-buffered

struct Ring {
  uint32_t stage = 0, phase = 0;     // raw-tile ring (producer / converters / MMA)
  uint32_t lo = 0, lo_phase = 0;     // lo-tile ring (converters / MMA)
  uint32_t acc_phase = 0;            // TMEM accumulator hand-off parity (MMA / epilogue), one tile in flight
};

struct Smem {
  uint8_t* tiles;        // 1024-byte aligned, (kStages + kLoStages) * kStageBytes
  uint64_t* full;        // [kStages] TMA bytes landed
  uint64_t* empty;       // [kStages] MMAs that read the raw stage have completed
  uint64_t* conv;        // [kLoStages] lo tiles written by the converter warps
  uint64_t* lo_empty;    // [kLoStages] MMAs that read the lo stage have completed
  uint64_t* tmem_full;   // [2]
  uint64_t* tmem_empty;  // [2]
  uint32_t* tmem_base;   // TMEM base address written by tcgen05.alloc
  double* red;           // [16] epilogue reduction scratch
  float* epi_stage;      // [8 warps][32][kEpiPitch] epilogue transpose tiles
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, Blackwell version field = 1 (layouts verified by tests/cuda/tc_selftest.cu).
//   K-major  tile: rows of 64 B (16 floats of K), SWIZZLE_64B, 8-row groups 512 B apart (SBO).
//   MN-major tile: chunks of [16 K-rows][128 B of MN]. 32-bit MN-major operands only exist in the
//                  SWIZZLE_128B_BASE32B layout (32-byte chunks XOR row%4; TMA mode SWIZZLE_128B_ATOM_32B):
//                  4-K-row groups 512 B apart (SBO), MN chunks 2048 B apart (LBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? ((uint32_t)kMnChunkBytes >> 4) : 1u;
  const uint64_t sbo = 512u >> 4;
  const uint64_t layout = mn_major ? 1ull : 4ull;   // SWIZZLE_128B_BASE32B : SWIZZLE_64B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

__device__ __forceinline__ void setup(Smem& sm, uint8_t* dyn_smem) {
  uintptr_t base = (reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~(uintptr_t)1023;
  sm.tiles = reinterpret_cast<uint8_t*>(base);
  uint8_t* tail = sm.tiles + (kStages + kLoStages) * kStageBytes;
  sm.full = reinterpret_cast<uint64_t*>(tail);
  sm.empty = sm.full + kStages;
  sm.conv = sm.empty + kStages;
  sm.lo_empty = sm.conv + kLoStages;
  sm.tmem_full = sm.lo_empty + kLoStages;
  sm.tmem_empty = sm.tmem_full + 2;
  sm.tmem_base = reinterpret_cast<uint32_t*>(sm.tmem_empty + 2);
  sm.red = reinterpret_cast<double*>(tail + 256);          // 16 doubles
  sm.epi_stage = reinterpret_cast<float*>(tail + 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < kLoStages; ++i) { mbar_init(&sm.conv[i], kConvWarps); mbar_init(&sm.lo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x / 32 == 2) tmem_alloc(sm.tmem_base, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
}

__device__ __forceinline__ void teardown(Smem& sm) {
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x / 32 == 2) tmem_dealloc(*sm.tmem_base, kTmemCols);
}

// Timeline trace (debug): CTA 0 appends (tag, clock) pairs per role; 4 roles x 1024 entries.
__device__ __forceinline__ void trace(const DevState* S, int role, int& n, int tag) {
  if (S->dbg && blockIdx.x == 0 && n < 511) {
    S->dbg[role * 1024 + 2 * n] = tag;
    S->dbg[role * 1024 + 2 * n + 1] = clock64();
    ++n;
  }
}

// Decode the work item of a GEMM phase (same arithmetic in every role).
struct Item { const GemmDesc* g; int chain, tile_m, tile_n; };

__device__ __forceinline__ bool decode_item(const DevState* S, const PhaseDesc& ph, int c_begin, int item, Item& it) {
  const GemmDesc& ga = S->gemm[ph.g0];
  const int ta = ga.tm * ga.tn;
  const int tb = (ph.g1 >= 0) ? S->gemm[ph.g1].tm * S->gemm[ph.g1].tn : 0;
  const int per_chain = ta + tb;
  it.chain = c_begin + item / per_chain;
  int t = item % per_chain;
  it.g = (t < ta) ? &ga : &S->gemm[ph.g1];
  if (t >= ta) t -= ta;
  it.tile_m = t / it.g->tn;
  it.tile_n = t % it.g->tn;
  return !(it.g->needLang && !S->ch.is_lang[it.chain]);
}

__device__ __forceinline__ int phase_items(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end) {
  const GemmDesc& ga = S->gemm[ph.g0];
  const int ta = ga.tm * ga.tn;
  const int tb = (ph.g1 >= 0) ? S->gemm[ph.g1].tm * S->gemm[ph.g1].tn : 0;
  return (ta + tb) * (c_end - c_begin);
}

// ---------------------------------------------------------------------------------------------------
// One GEMM phase on the tensor cores. All 256 threads call; returns after every tile of this CTA is stored.
// ---------------------------------------------------------------------------------------------------
__device__ void gemm_phase(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, Smem& sm, Ring& ring) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = phase_items(S, ph, c_begin, c_end);
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(S->tmaps);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int ntr = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        Item it;
        if (!decode_item(S, ph, c_begin, item, it)) continue;
        const GemmDesc& g = *it.g;
        const int nkb = (g.K + kBK - 1) / kBK;
        const int m0 = it.tile_m * kBM, n0 = it.tile_n * g.tcBN;
        const int ca = g.sA ? it.chain : 0, cb = g.sB ? it.chain : 0;
        const uint32_t bytes = (uint32_t)kATileBytes + (uint32_t)g.tcBBytes;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty[ring.stage], ring.phase ^ 1);
          trace(S, 0, ntr, kb);
          uint8_t* st = sm.tiles + ring.stage * kStageBytes;
          mbar_expect_tx(&sm.full[ring.stage], bytes);
          const int k0 = kb * kBK;
          if (g.aMode == 0) {
            tma_load_3d(st, &maps[g.tmIdx[0]], &sm.full[ring.stage], k0, m0, ca);
          } else {
            // MN-major: one [16 K-rows][32 floats] box per 32-wide chunk of the tile dimension
#pragma unroll
            for (int c = 0; c < kBM / 32; ++c)
              tma_load_3d(st + c * kMnChunkBytes, &maps[g.tmIdx[0]], &sm.full[ring.stage], m0 + c * 32, k0, ca);
          }
          uint8_t* sb = st + kATileBytes;
          if (g.bMode == 0) {
            tma_load_3d(sb, &maps[g.tmIdx[1]], &sm.full[ring.stage], k0, n0, cb);
          } else {
            const int nch = g.tcBN / 32;
            for (int c = 0; c < nch; ++c)
              tma_load_3d(sb + c * kMnChunkBytes, &maps[g.tmIdx[1]], &sm.full[ring.stage], n0 + c * 32, k0, cb);
          }
          if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int ntr = 0;
      const uint32_t tmem = *sm.tmem_base;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        Item it;
        if (!decode_item(S, ph, c_begin, item, it)) continue;
        const GemmDesc& g = *it.g;
        const int nkb = (g.K + kBK - 1) / kBK;
        mbar_wait(&sm.tmem_empty[0], ring.acc_phase ^ 1);
        fence_after_sync();
        // The tensor core adds into its FP32 accumulator with truncation, a bias that grows with the number of
        // accumulating MMAs. Keep the chains short: the two small cross terms go to their own accumulator and the
        // hi*hi term rotates over tcMain accumulators; the epilogue sums them in FP32 with round-to-nearest.
        const uint32_t d_cross = tmem + (uint32_t)(g.tcMain * g.tcBN);
        const uint32_t astep = g.aMode ? 1024u : 32u, bstep = g.bMode ? 1024u : 32u;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[ring.stage], ring.phase);
          mbar_wait(&sm.conv[ring.lo], ring.lo_phase);
          trace(S, 1, ntr, kb);
          fence_after_sync();
          const uint32_t sa = smem_u32(sm.tiles + ring.stage * kStageBytes);
          const uint32_t sb = sa + kATileBytes;
          const uint32_t la = smem_u32(sm.tiles + (kStages + ring.lo) * kStageBytes);
          const uint32_t lb = la + kATileBytes;
          const int nkk = min(kBK / 8, (g.K - kb * kBK + 7) / 8);
          for (int kk = 0; kk < nkk; ++kk) {
            const uint64_t a_hi = make_desc(sa + kk * astep, g.aMode != 0);
            const uint64_t a_lo = make_desc(la + kk * astep, g.aMode != 0);
            const uint64_t b_hi = make_desc(sb + kk * bstep, g.bMode != 0);
            const uint64_t b_lo = make_desc(lb + kk * bstep, g.bMode != 0);
            const uint32_t d_main = tmem + (uint32_t)((kb % g.tcMain) * g.tcBN);
            umma_tf32(d_cross, a_lo, b_hi, g.idesc, (kb | kk) ? 1u : 0u);
            umma_tf32(d_cross, a_hi, b_lo, g.idesc, 1u);
            umma_tf32(d_main, a_hi, b_hi, g.idesc, (kb >= g.tcMain || kk) ? 1u : 0u);
          }
          umma_commit(&sm.empty[ring.stage]);   // frees the raw slot when these MMAs have read it
          umma_commit(&sm.lo_empty[ring.lo]);   // ... and the lo slot
          if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
          if (++ring.lo == kLoStages) { ring.lo = 0; ring.lo_phase ^= 1; }
        }
        trace(S, 1, ntr, 1000);
        umma_commit(&sm.tmem_full[0]);   // accumulators complete -> epilogue
        ring.acc_phase ^= 1;
      }
    }
  } else {
    // ================= warps 2..11: hi/lo conversion of every K block of the tile, then (warps 4..11) the epilogue ====
    // lo = x - trunc_tf32(x), element-wise on the swizzled tile bytes (position independent)
    const int ct = threadIdx.x - 64;   // 0..319
    constexpr int kConvThreads = kConvWarps * 32;
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int eg = (warp - 4) >> 2;       // 0/1: which alternating 32-column chunks this warp drains
    const int ew = warp - 4;              // 0..7 (epilogue warps)
    const int et = threadIdx.x - 128;     // 0..255 (epilogue threads)
    const uint32_t tmem = *sm.tmem_base;
    int ntr = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      Item it;
      if (!decode_item(S, ph, c_begin, item, it)) continue;
      const GemmDesc& g = *it.g;
      {
        const int nkb = (g.K + kBK - 1) / kBK;
        const int nB4 = g.tcBBytes / 16;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.full[ring.stage], ring.phase);
          mbar_wait(&sm.lo_empty[ring.lo], ring.lo_phase ^ 1);
          if (ct == 0) trace(S, 2, ntr, kb);
          const uint8_t* st = sm.tiles + ring.stage * kStageBytes;
          uint8_t* lt = sm.tiles + (kStages + ring.lo) * kStageBytes;
          const float4* ar = reinterpret_cast<const float4*>(st);
          float4* al = reinterpret_cast<float4*>(lt);
          const float4* br = reinterpret_cast<const float4*>(st + kATileBytes);
          float4* bl = reinterpret_cast<float4*>(lt + kATileBytes);
#pragma unroll
          for (int i = ct; i < kATileBytes / 16; i += kConvThreads) {
            const float4 v = ar[i];
            float4 l, h;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            al[i] = l;
          }
#pragma unroll 3
          for (int i = ct; i < nB4; i += kConvThreads) {
            const float4 v = br[i];
            float4 l, h;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            bl[i] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> tensor-core (async proxy) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.conv[ring.lo]);
          if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
          if (++ring.lo == kLoStages) { ring.lo = 0; ring.lo_phase ^= 1; }
        }
      }
      if (warp < 4) continue;
      // ---------------- epilogue: TMEM -> registers -> fused epilogue -> HBM ----------------
      const int chain = it.chain;
      const int m0 = it.tile_m * kBM, n0 = it.tile_n * g.tcBN;
      const int r = m0 + q * 32 + lane;
      float* C = g.C ? g.C + (size_t)chain * g.sC : nullptr;
      const float* bias = g.bias ? g.bias + (size_t)chain * g.sBias : nullptr;
      const float* aux = g.aux ? g.aux + (size_t)chain * g.sAux : nullptr;
      float* gbias = g.gbias ? g.gbias + (size_t)chain * g.sGb : nullptr;
      mbar_wait(&sm.tmem_full[0], ring.acc_phase);
      if (et == 0) trace(S, 3, ntr, 0);
      fence_after_sync();
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
      const int nacc = min(g.tcMain, (g.K + kBK - 1) / kBK);   // main accumulators that were written
      float l_sse = 0.f, l_sd = 0.f;
      const int n_end = min(g.N, n0 + g.tcBN);   // this tile owns output columns [n0, n_end)
      const int ncols = n_end - n0;
      float* stg = sm.epi_stage + ew * (32 * kEpiPitch);   // this warp's 32 x 32 staging tile (pitch 36 floats)
      const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;  // coalesced view: 4 rows x 128 B per instruction
      const int r0 = m0 + q * 32;
      // aux tile loads (sigmoid outputs / data) are coalesced (4 rows x 128 B per instruction) and issued one
      // chunk ahead so that their L2 latency overlaps the TMEM drain and the math of the current chunk
      float4 t[8];
      auto load_aux = [&](int c0n) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = r0 + i * 4 + sub_r, cc = n0 + c0n + sub_c;
          t[i] = (c0n < ncols && rr < g.M && cc < n_end) ? *reinterpret_cast<const float4*>(&aux[(size_t)rr * g.ldaux + cc])
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (aux) load_aux(eg * 32);
      for (int c0 = eg * 32; c0 < ncols; c0 += 64) {
        uint32_t raw[32], part[32];
        tmem_ld32(tbase + c0, raw);
        for (int a = 1; a <= nacc; ++a) {   // a == nacc: the cross-term accumulator
          tmem_ld32(tbase + (uint32_t)((a == nacc ? g.tcMain : a) * g.tcBN) + c0, part);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(part[j]));
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const int cbase = n0 + c0;
        if (et == 0) trace(S, 3, ntr, 10);
        // ---- aux tile (sigmoid outputs / data), loaded coalesced and transposed through shared memory
        if (aux) {
#pragma unroll
          for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&stg[(i * 4 + sub_r) * kEpiPitch + sub_c]) = t[i];
          __syncwarp();
          load_aux(c0 + 64);   // prefetch the next chunk of this warp
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 a4 = *reinterpret_cast<const float4*>(&stg[lane * kEpiPitch + j4 * 4]);
            if (g.epi == EPI_LOSS) {
              part[j4 * 4 + 0] = __float_as_uint(a4.x); part[j4 * 4 + 1] = __float_as_uint(a4.y);
              part[j4 * 4 + 2] = __float_as_uint(a4.z); part[j4 * 4 + 3] = __float_as_uint(a4.w);
            } else {   // EPI_DSIG
              v[j4 * 4 + 0] *= a4.x * (1.0f - a4.x); v[j4 * 4 + 1] *= a4.y * (1.0f - a4.y);
              v[j4 * 4 + 2] *= a4.z * (1.0f - a4.z); v[j4 * 4 + 3] *= a4.w * (1.0f - a4.w);
            }
          }
          __syncwarp();
        }
        if (et == 0) trace(S, 3, ntr, 11);
        if (bias) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            if (cbase + j4 * 4 < n_end) {
              const float4 b4 = *reinterpret_cast<const float4*>(&bias[cbase + j4 * 4]);   // warp-uniform address
              v[j4 * 4 + 0] += b4.x; v[j4 * 4 + 1] += b4.y; v[j4 * 4 + 2] += b4.z; v[j4 * 4 + 3] += b4.w;
            }
          }
        }
        if (g.epi == EPI_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
        } else if (g.epi == EPI_LOSS) {
          const bool rv = (r < g.M);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dd = v[j] - __uint_as_float(part[j]);
            if (rv && cbase + j < n_end) { l_sse = fmaf(dd, dd, l_sse); l_sd += dd; }
            v[j] = dd * g.scale;
          }
        }
        if (et == 0) trace(S, 3, ntr, 12);
        // ---- transpose through shared memory and store coalesced
        if (C) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(&stg[lane * kEpiPitch + j4 * 4]) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
          __syncwarp();
          const int lim = (g.epi == EPI_GRADW && g.biasCol >= 0) ? min(n_end, g.biasCol) : n_end;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = r0 + i * 4 + sub_r, cc = cbase + sub_c;
            if (rr >= g.M || cc >= n_end) continue;
            const float4 o = *reinterpret_cast<const float4*>(&stg[(i * 4 + sub_r) * kEpiPitch + sub_c]);
            float* dst = &C[(size_t)rr * g.ldc + cc];
            if (cc + 4 <= lim) {
              *reinterpret_cast<float4*>(dst) = o;
            } else {
              const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (cc + j < lim) dst[j] = ov[j];
                else if (g.epi == EPI_GRADW && cc + j == g.biasCol) gbias[rr] = ov[j];   // ones-column trick
              }
            }
          }
          __syncwarp();
        }
      }
      // release the accumulator stage
      fence_before_sync();
      __syncwarp();
      if (et == 0) trace(S, 3, ntr, 1);
      if (lane == 0) mbar_arrive(&sm.tmem_empty[0]);
      ring.acc_phase ^= 1;
      if (g.epi == EPI_LOSS) {
        double s0 = (double)l_sse, s1 = (double)l_sd;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (lane == 0) { sm.red[ew * 2] = s0; sm.red[ew * 2 + 1] = s1; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          double* p = g.part + ((size_t)chain * g.partStride + (size_t)it.tile_m * g.tn + it.tile_n) * 2;
          double a0 = 0.0, a1 = 0.0;
          for (int w = 0; w < kEpiWarps; ++w) { a0 += sm.red[w * 2]; a1 += sm.red[w * 2 + 1]; }
          p[0] = a0;
          p[1] = a1;
        }
      }
    }
  }
  // generic-proxy stores of this phase must be ordered before later TMA (async-proxy) reads
  fence_proxy_async();
  __syncthreads();
}

}  // namespace tc
}  // namespace bae
