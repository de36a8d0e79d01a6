// tcgen05 / TMEM / TMA tile engine (sm_100a): chain-batched GEMM tiles with 3xTF32 split operands.
//
//   D[128 x BN] (FP32, TMEM) += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        per 8-wide K step
//
// Operands live in HBM as plain FP32. The tensor core reads a 32-bit container and ignores the low 13
// mantissa bits (measured: tests/cuda/tc_selftest.cu), so the raw tile IS the `hi` operand; two converter
// warps compute lo = x - trunc_tf32(x) (exact in FP32) into a second shared-memory tile. Three kind::tf32
// MMAs per K step then recover ~FP32 products with FP32 accumulation in TMEM. Raw tiles are staged by TMA
// (128-byte swizzle) into a 2-stage shared-memory ring; one elected thread issues tcgen05.mma; four epilogue
// warps drain TMEM with tcgen05.ld and apply the fused epilogue (bias / sigmoid / loss / sigmoid-derivative /
// gradient routing).
//
// Warp roles inside the 256-thread CTA: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-3 = hi/lo
// converters, warps 4-7 = epilogue.
#pragma once
#include <cuda.h>

#include "bae_common.cuh"

namespace bae {
namespace tc {

constexpr int kStages = 2;
constexpr int kBM = 128;
constexpr int kBK = 32;                        // floats per K block = one 128-byte swizzle row
constexpr int kATileBytes = kBM * kBK * 4;     // 16 KB
constexpr int kBTileBytes = 256 * kBK * 4;     // 32 KB (BN <= 256)
constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;  // A_hi, A_lo, B_hi, B_lo = 96 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
// stage layout: [A raw 16 KB][A lo 16 KB][B raw 32 KB][B lo 32 KB]
constexpr int kTmemCols = 512;                 // (tcMain + 1) accumulators of BN columns each, single-buffered

struct Ring {
  uint32_t stage = 0, phase = 0;     // shared-memory ring (producer / MMA)
  uint32_t acc_phase = 0;            // TMEM accumulator hand-off parity (MMA / epilogue), one tile in flight
};

struct Smem {
  uint8_t* tiles;        // 1024-byte aligned, kStages * kStageBytes
  uint64_t* full;        // [kStages] TMA bytes landed
  uint64_t* conv;        // [kStages] lo tiles written by the converter warps
  uint64_t* empty;       // [kStages] MMAs that read the stage have completed
  uint64_t* tmem_full;   // [2]
  uint64_t* tmem_empty;  // [2]
  uint32_t* tmem_base;   // TMEM base address written by tcgen05.alloc
  double* red;           // [8] epilogue reduction scratch
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

// Shared-memory matrix descriptor, Blackwell version field = 1.
//   K-major  tile: rows of 128 B (32 floats of K), SWIZZLE_128B (16-byte chunks), 8-row groups 1024 B apart (SBO).
//   MN-major tile: chunks of [32 K-rows][128 B of MN]. 32-bit MN-major operands only exist in the
//                  SWIZZLE_128B_BASE32B layout (32-byte chunks XOR row%4; TMA mode SWIZZLE_128B_ATOM_32B):
//                  4-K-row groups 512 B apart (SBO), MN chunks 4096 B apart (LBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  const uint64_t lbo = mn_major ? (4096u >> 4) : 1u;
  const uint64_t sbo = mn_major ? (512u >> 4) : (1024u >> 4);
  const uint64_t layout = mn_major ? 1ull : 2ull;   // SWIZZLE_128B_BASE32B : SWIZZLE_128B
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

__device__ __forceinline__ void setup(Smem& sm, uint8_t* dyn_smem) {
  uintptr_t base = (reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~(uintptr_t)1023;
  sm.tiles = reinterpret_cast<uint8_t*>(base);
  uint8_t* tail = sm.tiles + kStages * kStageBytes;
  sm.full = reinterpret_cast<uint64_t*>(tail);
  sm.conv = sm.full + kStages;
  sm.empty = sm.conv + kStages;
  sm.tmem_full = sm.empty + kStages;
  sm.tmem_empty = sm.tmem_full + 2;
  sm.tmem_base = reinterpret_cast<uint32_t*>(sm.tmem_empty + 2);
  sm.red = reinterpret_cast<double*>(tail + 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.conv[i], 2); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 4); }
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
          uint8_t* st = sm.tiles + ring.stage * kStageBytes;
          mbar_expect_tx(&sm.full[ring.stage], bytes);
          const int k0 = kb * kBK;
          if (g.aMode == 0) {
            tma_load_3d(st, &maps[g.tmIdx[0]], &sm.full[ring.stage], k0, m0, ca);
          } else {
            // MN-major: one [32 K-rows][32 floats] box per 32-wide chunk of the tile dimension
#pragma unroll
            for (int c = 0; c < kBM / 32; ++c)
              tma_load_3d(st + c * 4096, &maps[g.tmIdx[0]], &sm.full[ring.stage], m0 + c * 32, k0, ca);
          }
          uint8_t* sb = st + 2 * kATileBytes;
          if (g.bMode == 0) {
            tma_load_3d(sb, &maps[g.tmIdx[1]], &sm.full[ring.stage], k0, n0, cb);
          } else {
            const int nch = g.tcBN / 32;
            for (int c = 0; c < nch; ++c)
              tma_load_3d(sb + c * 4096, &maps[g.tmIdx[1]], &sm.full[ring.stage], n0 + c * 32, k0, cb);
          }
          if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
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
          mbar_wait(&sm.conv[ring.stage], ring.phase);
          fence_after_sync();
          const uint32_t sa = smem_u32(sm.tiles + ring.stage * kStageBytes);
          const uint32_t sb = sa + 2 * kATileBytes;
          const int nkk = min(4, (g.K - kb * kBK + 7) / 8);
          for (int kk = 0; kk < nkk; ++kk) {
            const uint64_t a_hi = make_desc(sa + kk * astep, g.aMode != 0);
            const uint64_t a_lo = make_desc(sa + kATileBytes + kk * astep, g.aMode != 0);
            const uint64_t b_hi = make_desc(sb + kk * bstep, g.bMode != 0);
            const uint64_t b_lo = make_desc(sb + kBTileBytes + kk * bstep, g.bMode != 0);
            const uint32_t d_main = tmem + (uint32_t)((kb % g.tcMain) * g.tcBN);
            umma_tf32(d_cross, a_lo, b_hi, g.idesc, (kb | kk) ? 1u : 0u);
            umma_tf32(d_cross, a_hi, b_lo, g.idesc, 1u);
            umma_tf32(d_main, a_hi, b_hi, g.idesc, (kb >= g.tcMain || kk) ? 1u : 0u);
          }
          umma_commit(&sm.empty[ring.stage]);   // frees the smem slot when these MMAs have read it
          if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
        }
        umma_commit(&sm.tmem_full[0]);   // accumulators complete -> epilogue
        ring.acc_phase ^= 1;
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================= hi/lo converters: lo = x - trunc_tf32(x), element-wise on the swizzled tile bytes =================
    const int ct = threadIdx.x - 64;   // 0..63
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      Item it;
      if (!decode_item(S, ph, c_begin, item, it)) continue;
      const GemmDesc& g = *it.g;
      const int nkb = (g.K + kBK - 1) / kBK;
      const int nB4 = g.tcBBytes / 16;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&sm.full[ring.stage], ring.phase);
        uint8_t* st = sm.tiles + ring.stage * kStageBytes;
        const float4* ar = reinterpret_cast<const float4*>(st);
        float4* al = reinterpret_cast<float4*>(st + kATileBytes);
        const float4* br = reinterpret_cast<const float4*>(st + 2 * kATileBytes);
        float4* bl = reinterpret_cast<float4*>(st + 2 * kATileBytes + kBTileBytes);
#pragma unroll 8
        for (int i = ct; i < kATileBytes / 16; i += 64) {
          const float4 v = ar[i];
          float4 l, h;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          al[i] = l;
        }
#pragma unroll 8
        for (int i = ct; i < nB4; i += 64) {
          const float4 v = br[i];
          float4 l, h;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          bl[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> tensor-core (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.conv[ring.stage]);
        if (++ring.stage == kStages) { ring.stage = 0; ring.phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> fused epilogue -> HBM (hi, lo) =================
    const int q = warp & 3;
    const int et = threadIdx.x - 128;   // 0..127
    const uint32_t tmem = *sm.tmem_base;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      Item it;
      if (!decode_item(S, ph, c_begin, item, it)) continue;
      const GemmDesc& g = *it.g;
      const int chain = it.chain;
      const int m0 = it.tile_m * kBM, n0 = it.tile_n * g.tcBN;
      const int r = m0 + q * 32 + lane;
      float* C = g.C ? g.C + (size_t)chain * g.sC : nullptr;
      const float* bias = g.bias ? g.bias + (size_t)chain * g.sBias : nullptr;
      const float* aux = g.aux ? g.aux + (size_t)chain * g.sAux : nullptr;
      float* gbias = g.gbias ? g.gbias + (size_t)chain * g.sGb : nullptr;
      mbar_wait(&sm.tmem_full[0], ring.acc_phase);
      fence_after_sync();
      const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
      const int nacc = min(g.tcMain, (g.K + kBK - 1) / kBK);   // main accumulators that were written
      float l_sse = 0.f, l_sd = 0.f;
      const int n_end = min(g.N, n0 + g.tcBN);   // this tile owns output columns [n0, n_end)
      const int ncols = n_end - n0;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t raw[32], part[32];
        tmem_ld32(tbase + c0, raw);
        for (int a = 1; a <= nacc; ++a) {   // a == nacc: the cross-term accumulator
          tmem_ld32(tbase + (uint32_t)((a == nacc ? g.tcMain : a) * g.tcBN) + c0, part);
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(part[j]));
        }
        if (r < g.M) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int c = n0 + c0 + j4 * 4;
            const int nval = min(4, n_end - c);
            if (nval <= 0) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(raw[j4 * 4 + j]);
            float bv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
            if (bias) {
              float4 t = *reinterpret_cast<const float4*>(&bias[c]);
              bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
            }
            if (aux) {
              float4 t = *reinterpret_cast<const float4*>(&aux[(size_t)r * g.ldaux + c]);
              av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
            }
            switch (g.epi) {
              case EPI_SIGMOID:
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = 1.0f / (1.0f + expf(-(v[j] + bv[j])));
                break;
              case EPI_BIAS:
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = v[j] + bv[j];
                break;
              case EPI_LOSS:
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float dd = (v[j] + bv[j]) - av[j];
                  if (j < nval) { l_sse = fmaf(dd, dd, l_sse); l_sd += dd; }
                  v[j] = dd * g.scale;
                }
                break;
              case EPI_DSIG:
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = v[j] * av[j] * (1.0f - av[j]);
                break;
              default:
                break;
            }
            if (g.epi == EPI_GRADW && g.biasCol >= 0) {
              // the ones column appended to the activation operand makes column biasCol the bias gradient
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (c + j == g.biasCol) gbias[r] = v[j];
                else if (c + j < g.biasCol) C[(size_t)r * g.ldc + c + j] = v[j];
              }
              continue;
            }
            if (!C) continue;
            float* dst = &C[(size_t)r * g.ldc + c];
            if (nval == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else for (int j = 0; j < nval; ++j) dst[j] = v[j];
          }
        }
      }
      // release the accumulator stage
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmem_empty[0]);
      ring.acc_phase ^= 1;
      if (g.epi == EPI_LOSS) {
        double s0 = (double)l_sse, s1 = (double)l_sd;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(0xffffffffu, s0, o);
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (lane == 0) { sm.red[q * 2] = s0; sm.red[q * 2 + 1] = s1; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          double* p = g.part + ((size_t)chain * g.partStride + (size_t)it.tile_m * g.tn + it.tile_n) * 2;
          p[0] = sm.red[0] + sm.red[2] + sm.red[4] + sm.red[6];
          p[1] = sm.red[1] + sm.red[3] + sm.red[5] + sm.red[7];
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
