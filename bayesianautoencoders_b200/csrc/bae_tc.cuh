// tcgen05 / TMEM / TMA tile engine (sm_100a): chain-batched GEMM tiles with three-way split FP32 operands.
//
//   D[128 x BN] (FP32, TMEM) = sum_k  A_hi*B_hi   (accumulator 0)   +   A_lo*B_hi + A_hi*B_lo   (accumulator 1)
//
// Two engines share this file. The production one is the 3xFP16 split (kind::f16, scaled operands: further down, "3xFP16
// split"), with operands that are known before the GEMM runs arriving as pre-split images ("operand images") and the front
// half compiled per mode (FrontMode). The round-1 engine, kept as a fallback (BAE_TC_TF32=1, the persistent kernel), is
// the 3xTF32 split described first:
// Operands live in HBM as plain FP32. The tensor core reads a 32-bit container and ignores the low 13
// mantissa bits (measured: tests/cuda/tc_selftest.cu), so the raw tile IS the `hi` operand; converter warps
// compute lo = x - trunc_tf32(x) (exact in FP32) into a second shared-memory tile. Three kind::tf32 MMAs per
// 8-wide K step then recover ~FP32 products with FP32 accumulation in TMEM.
//
// CTA PAIRS (tcgen05 cta_group::2, the kernel is launched with a cluster dimension of 2): a pair computes a 256 x BN tile.
// CTA r of the pair loads and converts rows [128 r, 128 r + 128) of A and columns [BN/2 r, BN/2 r + BN/2) of B into ITS OWN
// shared memory; the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and
// writes each CTA's 128 accumulator rows into that CTA's TMEM; every CTA drains and post-processes its own rows. Per CTA
// and K block this loads and converts 16 KB instead of 23.5 KB for the same tensor-pipe time: the hi/lo conversion
// (shared-memory latency bound, tests/cuda/tc_microbench.cu) is what paces the tile loop.
//
// 512-thread CTA, one per SM, four warpgroups with their own register budgets (setmaxnreg):
//   warp 0      TMA producer: raw tiles (16 floats of K per stage; operand-image tiles by plain bulk copies) into a 7-stage ring
//   warp 1      MMA issuer (leader CTA only; warp-uniform loop, one elected lane issues)
//   warps 2-3, 12-15   hi/lo converters, three groups of two warps that take K blocks round-robin, so the
//                      wait -> LDS -> STS -> proxy fence -> arrive latency of one block overlaps the next ones
//   warps 4-11  epilogue: drain both accumulators into registers (tcgen05.ld), hand TMEM back to the MMA
//               issuer at once, then run the fused epilogue (bias / sigmoid / loss / sigmoid' / gradient
//               routing) from registers while the next tile's MMAs run. Values are transposed through
//               shared memory so every global load and store is a coalesced 128-byte row segment.
// Barriers live in every CTA at the same offsets. `full` (TMA landed) is local; `conv` (lo tiles written) and `tmem_empty`
// (accumulators drained) are used in the leader's copy and collect arrivals from both CTAs (remote arrives through mapa);
// `empty` (stage free) and `tmem_full` (accumulators complete) are signalled in both CTAs by multicast tcgen05.commit.
// The tensor core adds into its FP32 accumulator with truncation, a bias that grows with the number of
// accumulating MMAs: the two small cross terms therefore go to their own accumulator, and long contractions
// (weight gradients, K = rows) are cut into segments that one CTA pair processes back to back, adding each into the
// gradient with a TMA reduce-add (GemmDesc::ksplit).
#pragma once
#include <cuda.h>

#include "bae_common.cuh"

// Timing-experiment switches (skip the hi/lo conversion, the cross-term MMAs, the epilogue math, the A / B loads) exist
// only in builds with -DBAE_TC_DEBUG (tools/tc_experiments.sh): results are wrong when one is set, so the shipped library
// does not contain the branches and bae_create refuses BAE_TC_DBGFLAGS.
#ifdef BAE_TC_DEBUG
#define BAE_DBGF(S) ((S)->dbg_flags)
#else
#define BAE_DBGF(S) 0
#endif

namespace bae {
namespace tc {

constexpr int kThreadsTc = 512;
#ifndef BAE_TC_STAGES
#define BAE_TC_STAGES 7
#define BAE_TC_LOSTAGES 5
#endif
constexpr int kStages = BAE_TC_STAGES;         // raw tile pairs (TMA ring): deep, the loads come from HBM (~2000 cycles under load)
constexpr int kLoStages = BAE_TC_LOSTAGES;     // converted tiles (converter ring): three groups converting + two being read / waiting
constexpr int kBM = 128;
constexpr int kBK = 16;                        // floats of K per stage (64-byte rows: SWIZZLE_64B for K-major tiles)
constexpr int kMaxBN = 256;                    // pair tile width; each CTA holds BN/2 columns of B; 2 accumulators x 256 TMEM columns
constexpr int kATileBytes = kBM * kBK * 4;     // 8 KB
constexpr int kBTileBytes = (kMaxBN / 2) * kBK * 4;  // 8 KB: this CTA's half of the B tile
constexpr int kMnChunkBytes = kBK * 128;       // MN-major chunk: [16 K-rows][32 floats]
constexpr int kStageBytes = kATileBytes + kBTileBytes;          // one raw stage (A, B half) or one lo stage = 16 KB
constexpr int kEpiWarps = 8;
constexpr int kConvGroups = 3;                 // converter groups; K block k is converted by group k % 3
constexpr int kConvGroupWarps = 2;             // 64 threads: every tile pair is a multiple of 64 float4 (no predicates)
constexpr int kConvGroupThreads = kConvGroupWarps * 32;
constexpr int kEpiBytes = kEpiWarps * 4096;       // one dense, swizzled 32 x 32 staging tile per epilogue warp
constexpr int kSmemBytes = (kStages + kLoStages) * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/ + kEpiBytes;
static_assert(kSmemBytes + 1024 <= 232448, "dynamic + static shared memory of the tcgen05 kernels (static: reduction scratch, segment table)");
// tile-pair layout: [A 8 KB][B half 8 KB]; kStages raw pairs followed by kLoStages lo pairs
constexpr int kTmemCols = 512;                 // two accumulators of BN <= 256 columns each, single-buffered
constexpr int kRegsLeanPersist = 48, kRegsEpiPersist = 208;   // persistent kernel (3xTF32 only): its non-GEMM phases need the registers
#ifndef BAE_TC_REGS_LEAN
#define BAE_TC_REGS_LEAN 48
#define BAE_TC_REGS_EPI 208
#endif
constexpr int kRegsLean = BAE_TC_REGS_LEAN;    // producer / MMA / converter warpgroups
constexpr int kRegsEpi = BAE_TC_REGS_EPI;      // epilogue warpgroups (also run the non-GEMM phases)
static_assert(kRegsLean * 256 + kRegsEpi * 256 <= 65536, "register file");

struct Ring {
  uint32_t stage = 0, phase = 0;     // raw-tile ring (producer / converters / MMA)
  uint32_t lo = 0, lo_phase = 0;     // lo-tile ring (converters / MMA)
  uint32_t acc_phase = 0;            // TMEM accumulator hand-off parity (MMA / epilogue), one tile in flight
  uint32_t kbc = 0;                  // running K-block count of this CTA modulo kConvGroups
  __device__ __forceinline__ void advance() {
    if (++stage == kStages) { stage = 0; phase ^= 1; }
    if (++lo == kLoStages) { lo = 0; lo_phase ^= 1; }
  }
};

// Carve-up of the dynamic shared memory. Only the base pointer is kept in a register; every other address is the
// base plus a compile-time offset (the persistent kernel keeps this struct live across all phases, and nothing there
// may spill: with the carve-out at its maximum there is no L1 behind local memory).
struct Smem {
  uint8_t* tiles;        // 1024-byte aligned: tile rings, then the epilogue staging tiles, then barriers and scratch
  static constexpr int kEpiOff = (kStages + kLoStages) * kStageBytes;   // the Adam phases stage through [0, kEpiOff + kEpiBytes)
  static constexpr int kTail = kEpiOff + kEpiBytes;
  __device__ __forceinline__ float* epi_stage() const { return reinterpret_cast<float*>(tiles + kEpiOff); }           // [8 warps][32 rows x 128 B, swizzled]
  __device__ __forceinline__ uint64_t* full() const { return reinterpret_cast<uint64_t*>(tiles + kTail); }            // [kStages] TMA bytes landed
  __device__ __forceinline__ uint64_t* empty() const { return full() + kStages; }       // [kStages] MMAs that read the raw stage completed
  __device__ __forceinline__ uint64_t* conv() const { return empty() + kStages; }       // [kLoStages] lo pair written (leader: both CTAs)
  __device__ __forceinline__ uint64_t* lo_empty() const { return conv() + kLoStages; }  // [kLoStages] MMAs that read the lo stage completed
  __device__ __forceinline__ uint64_t* tmem_full() const { return lo_empty() + kLoStages; }  // [2]
  __device__ __forceinline__ uint64_t* tmem_empty() const { return tmem_full() + 2; }   // [2] (leader: both CTAs)
  __device__ __forceinline__ uint64_t* spare_bars() const { return tmem_empty() + 2; }  // [3] Adam staging
  __device__ __forceinline__ uint32_t* tmem_base() const { return reinterpret_cast<uint32_t*>(spare_bars() + 3); }   // TMEM base address
  __device__ __forceinline__ double* red() const { return reinterpret_cast<double*>(tiles + kTail + 256); }           // [16] reduction scratch
};
static_assert((2 * kStages + 2 * kLoStages + 4 + 3) * 8 + 4 <= 256, "barrier area");

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
  // try_wait with a large suspend-time hint: the hardware parks the thread until the phase flips instead of
  // letting sixteen warps spin on the barrier unit
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(a), "r"(parity), "r"(0x989680u)
      : "memory");
}

// One lane of a converged warp (warp-uniform code keeps descriptors and addresses in uniform registers; a
// `lane == 0` branch around the whole loop makes the compiler re-broadcast every operand of every UTMALDG / UTCHMMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- CTA-pair (cluster of 2) helpers
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the LEADER's copy of a shared-memory object (same offset in every CTA of the pair)
__device__ __forceinline__ uint32_t leader_addr(const void* p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(0u));
  return r;
}
__device__ __forceinline__ void mbar_arrive_leader(const void* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
// Signal only: a release at cluster scope costs ~1500 cycles per arrival in this kernel. The converters use this one: what
// they publish (the lo tile) stays in their own shared memory and is read by their own SM's tensor core, and the proxy
// fence before the arrival has already ordered the writes.
__device__ __forceinline__ void mbar_arrive_leader_relaxed(const void* bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(bar)) : "memory");
}
// wait that acquires at cluster scope: the other CTA's shared-memory / tensor-memory accesses are ordered before it
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAITC_DONE;\n\t"
      "bra WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t}"
      ::"r"(a), "r"(parity), "r"(0x989680u)
      : "memory");
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

// Plain (1-D) bulk copy global -> shared through the copy engine; bytes land on the same transaction barrier as the tensor loads
__device__ __forceinline__ void bulk_load(uint32_t dst_s, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_s), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
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

// Same load without the wait: the caller must execute a waiting load (tmem_ld32) before reading v.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
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
}

// Wait for every tcgen05.ld issued so far; `v` (the registers of the youngest load) is tied to the statement so that
// no consumer of those registers can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

// Register budgets per warpgroup (all four warps of a warpgroup execute the same instruction).
template <int kRegs> __device__ __forceinline__ void regs_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void regs_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

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

// Explicit shared-space accesses (a generic pointer into the dynamic shared-memory window compiles to LD.E / ST.E
// with a run-time address-space check).
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Global 16-byte load that stays where it is written (volatile asm): the epilogue issues its aux loads in two groups of
// four; left to the compiler they were all hoisted above the staging stores and then spilled (no L1 behind local memory).
__device__ __forceinline__ float4 ldg128_pinned(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// hi = what the tensor core reads from the raw FP32 container (it ignores the low 13 mantissa bits: truncation);
// lo = x - hi is exact in FP32 but has up to 24 significant bits, of which the tensor core would again keep only the top 11
// by TRUNCATION: x would then be represented with a one-sided error of up to 2^-21 |x| (mean 2^-22), a systematic
// shrink of every product that adds up over the layers. lo is therefore ROUNDED to the nearest TF32 value here (ties
// away from zero; an integer add on the bit pattern): the representation error of x becomes two-sided, zero-mean.
#ifndef BAE_LO_RN
#define BAE_LO_RN 1
#endif
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float l = x - hi;
#if BAE_LO_RN
  lo = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xFFFFE000u);
#else
  lo = l;
#endif
}

// ---------------------------------------------------------------------------------------------------
// 3xFP16 split (kind::f16 MMAs, K = 16 per instruction: half the tensor-pipe time of the 3xTF32 split).
// FP16 has the SAME 11 significant bits as TF32; what it lacks is exponent range (2^-14 .. 2^15 normal). Every operand is
// therefore multiplied by a power of two s = 2^e chosen from a bound on its magnitude (static for data and sigmoid outputs,
// tracked per chain for weights, BatchNorm outputs and deltas: DevState::amax) so that |x s| < 2^14, then split
//     hi = fp16_rn(x s)            11 significant bits (fewer only below 2^-14: 2^-27 of the operand's bound)
//     lo = fp16_rn(x s - hi)       the exact residual, again 11 bits (fewer only when |x s| < 2^-3, i.e. an absolute error
//                                  below 2^-38 of the operand's bound: far under the FP32 rounding of the sums it enters)
// and the three products hi*hi (accumulator 0), lo*hi + hi*lo (accumulator 1) are recombined in the epilogue as
// (acc0 + acc1) / (sA sB): powers of two throughout, so the result does not depend on the exponents chosen.
//
// The converter warps write the FP16 tiles in the NO-SWIZZLE canonical layout, which is the same byte formula for both
// majors: a "unit" = 8 consecutive elements along the operand's contiguous dimension = 16 bytes, unit index
//     U = 16 i + 8 j + r     at byte offset 16 U
//   K-major  (rows of K):   row = 8 i + r, k = 8 j .. 8 j + 7      (core matrix = 8 rows x 16 B; LBO = 128 B between the two
//   MN-major (rows of MN):  k = 8 j + r,  mn = 8 i .. 8 i + 7       K halves, SBO = 256 B between groups of 8 rows / 8 MN)
// A quarter warp writes 8 consecutive units = 128 contiguous bytes (conflict free), and its 8 reads of the TMA-written FP32
// tile fall into 8 different 16-byte bank groups in both source swizzles (for MN-major sources the two halves of a 32-byte
// swizzle unit are read in opposite order by the lanes with r >= 4).
constexpr int kF16TileBytes = 128 * kBK * 2;   // 4 KB: 128 rows (or MN) x 16 K of halves; a 16 KB lo stage = [A hi][A lo][B hi][B lo]
//
// OPERAND IMAGES (GemmDesc::aImg / bImg, DevState::preconv): an operand whose values and scale are known before the GEMM runs
// (the weights: rewritten once per proposal; the data matrices: once per upload) is split ONCE, by bae_wconv_kernel /
// bae_ximg_kernel, into an image in HBM that holds the units of the canonical layout in tile order:
//   K-major use  (contraction over the matrix's columns; row = 8 t8 + r, k = 16 kb + 8 j .. + 7):
//     unit (t8, kb, j, r)  at  kb * kbs + plane * (kbs / 2) + t8 * 256 + j * 128 + r * 16            (plane 0 = hi, 1 = lo)
//   MN-major use (contraction over the matrix's rows; k = 16 kb + 8 j + r, mn = 8 t8 .. + 7), T8 = kbs / 512 MN groups:
//     unit (t8, kb, j, r)  at  kb * kbs + plane * (kbs / 2) + j * (128 T8) + t8 * 128 + r * 16
// so the hi (lo) tile of `n` rows / MN from t0 (a multiple of 8) of K block kb is ONE contiguous run of 32 n bytes (K-major) or
// one run of 16 n bytes per K half (MN-major: tile layout [j][t8][r], i.e. descriptor strides LBO = 16 n, SBO = 128): the producer
// copies the runs into the operand's half of the raw stage (cp.async.bulk, no tensor map), the tensor core reads them there
// and the converter warps skip that operand - no LDS / STS / ALU work, 16 KB less shared-memory traffic per operand and K
// block. The values are those the converters would have produced (same scale exponent, same roundings): bit-identical results.
// ACTIVATION images (GemmDesc::aOut; MN-major use by the weight-gradient GEMMs): the forward GEMM that reads the activation as
// its K-major A operand converts every unit anyway; the converter group that finished a K block of a tile_n == 0 (1) item sends
// the converted [A hi] ([A lo]) tile to the image with one 4-D box store (the image's [j][t8][r] order is the tile's
// [row group][k half][r] order under a permutation of the dimensions, which the tensor map expresses).

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {   // a at the lower address
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void unpack_half2(uint32_t h, float& a, float& b) {
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(a), "=f"(b) : "r"(h));
}

// source addresses of unit U of an FP32 tile written by TMA; mode 0: K-major rows of 64 B, SWIZZLE_64B; mode 1: MN-major
// chunks of [16 k][32 mn], SWIZZLE_128B_ATOM_32B. Returns whether the two 16-byte halves come back swapped.
__device__ __forceinline__ bool f16_unit_src(uint32_t src_s, int U, int mode, uint32_t& a0, uint32_t& a1) {
  const int r = U & 7, j = (U >> 3) & 1, i = U >> 4;
  if (mode == 0) {
    const int row = i * 8 + r;
    const uint32_t sw = (uint32_t)(row >> 1) & 3u, base = src_s + (uint32_t)row * 64u;
    a0 = base + ((((uint32_t)(2 * j)) ^ sw) << 4);
    a1 = base + ((((uint32_t)(2 * j + 1)) ^ sw) << 4);
    return false;
  }
  const int k = j * 8 + r;
  const uint32_t base = src_s + (uint32_t)(i >> 2) * 2048u + (uint32_t)k * 128u + ((((uint32_t)i & 3u) ^ ((uint32_t)k & 3u)) << 5);
  const bool swap = (r & 4) != 0;
  a0 = base + (swap ? 16u : 0u);
  a1 = base + (swap ? 0u : 16u);
  return swap;
}

// 8 scaled values -> 16 B of hi halves and 16 B of lo halves
__device__ __forceinline__ void f16_split8(const float4& v0, const float4& v1, float s, uint4& hi, uint4& lo) {
  const float x[8] = {v0.x * s, v0.y * s, v0.z * s, v0.w * s, v1.x * s, v1.y * s, v1.z * s, v1.w * s};
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    h[p] = pack_half2(x[2 * p], x[2 * p + 1]);
    float f0, f1;
    unpack_half2(h[p], f0, f1);
    l[p] = pack_half2(x[2 * p] - f0, x[2 * p + 1] - f1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void sts128u(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Converts units [u0, u0 + stride, ...) < n_units of one operand tile (two units in flight per thread)
__device__ __forceinline__ void f16_convert_tile(uint32_t src_s, uint32_t hi_s, uint32_t lo_s, int mode, float s, int u0, int stride,
                                                 int n_units) {
  for (int U = u0; U < n_units; U += 2 * stride) {
    const int U2 = U + stride;
    const bool two = U2 < n_units;
    uint32_t a0, a1, b0 = 0, b1 = 0;
    const bool swa = f16_unit_src(src_s, U, mode, a0, a1);
    bool swb = false;
    if (two) swb = f16_unit_src(src_s, U2, mode, b0, b1);
    float4 va0 = lds128(a0), va1 = lds128(a1), vb0, vb1;
    if (two) { vb0 = lds128(b0); vb1 = lds128(b1); }
    uint4 hi, lo;
    if (swa) f16_split8(va1, va0, s, hi, lo); else f16_split8(va0, va1, s, hi, lo);
    sts128u(hi_s + (uint32_t)U * 16u, hi);
    sts128u(lo_s + (uint32_t)U * 16u, lo);
    if (two) {
      if (swb) f16_split8(vb1, vb0, s, hi, lo); else f16_split8(vb0, vb1, s, hi, lo);
      sts128u(hi_s + (uint32_t)U2 * 16u, hi);
      sts128u(lo_s + (uint32_t)U2 * 16u, lo);
    }
  }
}

// The same conversion for a 64-thread converter group of the engine (thread gt takes units gt, gt + 64, ...): r = U & 7 and
// j = (U >> 3) & 1 are then constants of the thread and i = (gt >> 4) + 4 n, which makes BOTH source swizzles collapse to
//     address(n) = tile + off(thread, mode) + 2048 n,      second half of the unit at address(n) ^ 16
// (K-major: (row >> 1) & 3 = (r >> 1) & 3 for every n; MN-major: the 32-column chunk index is n). Lanes with r >= 4 of an
// MN-major source start with the upper half (`swap`): their two loads come back in the opposite order.
struct ConvLane {
  uint32_t off[2];   // per source mode
  bool swap1;        // mode 1 only
};
__device__ __forceinline__ ConvLane conv_lane(int gt) {
  const uint32_t r = gt & 7, j = (gt >> 3) & 1, i0 = gt >> 4;
  ConvLane c;
  c.off[0] = (8 * i0 + r) * 64 + (((2 * j) ^ ((r >> 1) & 3)) << 4);
  const uint32_t k = 8 * j + r;
  // (the j term keeps the 8-byte stores of conv_units conflict free: units (r, 0) and (r, 1) are 128 B apart)
  c.swap1 = (((r >> 2) ^ j) & 1) != 0;
  c.off[1] = k * 128 + ((i0 ^ (k & 3)) << 5) + (c.swap1 ? 16u : 0u);
  return c;
}
__device__ __forceinline__ void sts64u(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
// 4 scaled values -> 8 B of hi halves and 8 B of lo halves
__device__ __forceinline__ void f16_split4(const float4& v, float s, uint32_t (&h)[2], uint32_t (&l)[2]) {
  const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    h[p] = pack_half2(x[2 * p], x[2 * p + 1]);
    float f0, f1;
    unpack_half2(h[p], f0, f1);
    l[p] = pack_half2(x[2 * p] - f0, x[2 * p + 1] - f1);
  }
}
// units gt + 64 n, n = 0 .. kN - 1, of one operand tile: all loads first (2 kN float4 in flight), then the splits.
// kSwap: the source is MN-major: lanes with `swap` set loaded the upper half of the unit first. Nothing is exchanged in
// registers; each half is stored as 8 bytes, at exchanged positions for those lanes (half-warp-wide conflict free, see conv_lane).
template <int kN, bool kSwap>
__device__ __forceinline__ void conv_units(uint32_t src_t, uint32_t hi_t, uint32_t lo_t, bool swap, float s) {
  float4 p[kN], q[kN];
#pragma unroll
  for (int u = 0; u < kN; ++u) {
    p[u] = lds128(src_t + (uint32_t)u * 2048u);
    q[u] = lds128((src_t + (uint32_t)u * 2048u) ^ 16u);
  }
  if (kSwap) {
    const uint32_t d8 = swap ? 8u : 0u;
#pragma unroll
    for (int u = 0; u < kN; ++u) {
      uint32_t h[2], l[2];
      f16_split4(p[u], s, h, l);
      sts64u((hi_t + (uint32_t)u * 1024u) + d8, h[0], h[1]);
      sts64u((lo_t + (uint32_t)u * 1024u) + d8, l[0], l[1]);
      f16_split4(q[u], s, h, l);
      sts64u(((hi_t + (uint32_t)u * 1024u) + d8) ^ 8u, h[0], h[1]);
      sts64u(((lo_t + (uint32_t)u * 1024u) + d8) ^ 8u, l[0], l[1]);
    }
  } else {
#pragma unroll
    for (int u = 0; u < kN; ++u) {
      uint4 hi, lo;
      f16_split8(p[u], q[u], s, hi, lo);
      sts128u(hi_t + (uint32_t)u * 1024u, hi);
      sts128u(lo_t + (uint32_t)u * 1024u, lo);
    }
  }
}
// up to four units (nv of them) of one operand tile
template <bool kSwap>
__device__ __forceinline__ void conv_tile(uint32_t src_t, uint32_t hi_t, uint32_t lo_t, bool swap, float s, int nv) {
  if (nv >= 2) conv_units<2, kSwap>(src_t, hi_t, lo_t, swap, s);
  else if (nv == 1) conv_units<1, kSwap>(src_t, hi_t, lo_t, swap, s);
  if (nv >= 4) conv_units<2, kSwap>(src_t + 4096, hi_t + 2048, lo_t + 2048, swap, s);
  else if (nv == 3) conv_units<1, kSwap>(src_t + 4096, hi_t + 2048, lo_t + 2048, swap, s);
}
// Shared-memory descriptor of an FP16 operand tile in the no-swizzle layout above (either major): LBO = 128 B, SBO = 256 B
__device__ __forceinline__ uint64_t make_desc_f16(uint32_t saddr, uint32_t lbo = 128u, uint32_t sbo = 256u) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// 4-D box store shared -> global (a converted A tile into an activation image, see "operand images")
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src_s, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src_s), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// f16: the raw (FP32) stages are read by this CTA's converter warps only, which release them themselves (one arrival per warp
// of the converting group); in the 3xTF32 engine the tensor core reads them too and the MMA issuer's commit releases them.
// f16: arrivals that free a raw stage / a converted slot (see gemm_phase_front):
//   n_empty    converter warps of the group (2), + 1 where stages may hold operand images (the MMAs' commit, or the group again)
//   n_lo_empty the MMAs' commit, + 1 in phases whose converted tiles are also read by image stores (the storing group)
__device__ __forceinline__ void setup(Smem& sm, uint8_t* dyn_smem, bool f16 = false, int n_empty = 1, int n_lo_empty = 1) {
  uintptr_t base = (reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~(uintptr_t)1023;
  sm.tiles = reinterpret_cast<uint8_t*>(base);
  if (threadIdx.x == 0) {
    // `conv` and `tmem_empty` collect arrivals from both CTAs of the pair (only the leader's copies are waited on)
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.full()[i], 1); mbar_init(&sm.empty()[i], f16 ? n_empty : 1); }
    for (int i = 0; i < kLoStages; ++i) { mbar_init(&sm.conv()[i], 2 * kConvGroupWarps); mbar_init(&sm.lo_empty()[i], f16 ? n_lo_empty : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full()[i], 1); mbar_init(&sm.tmem_empty()[i], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x / 32 == 2) tmem_alloc(sm.tmem_base(), kTmemCols);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything can arrive on them
  fence_after_sync();
}

__device__ __forceinline__ void teardown(Smem& sm) {
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // both CTAs are done with the pair's tensor memory and with each other's shared memory
  if (threadIdx.x / 32 == 2) tmem_dealloc(*sm.tmem_base(), kTmemCols);
}

// Timeline trace (debug): CTA 0 appends (tag, clock) pairs per role; 4 roles x 1024 entries.
__device__ __forceinline__ void trace(long long* dbg, int role, int& n, int tag) {
  if (dbg && n < 511) {
    dbg[role * 1024 + 2 * n] = tag;
    dbg[role * 1024 + 2 * n + 1] = clock64();
    ++n;
  }
}

// Store of one 32 x 32 box from a 128-byte-swizzled staging tile (shared -> global through the copy engine)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src_s, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src_s), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// Same box, ADDED to global memory (FP32 add performed at the L2; one box per instruction, elements do not interact)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* tm, uint32_t src_s, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src_s), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Work items of a GEMM phase (same arithmetic in every role): one item = one output tile of one chain. The K SEGMENTS of
// a weight-gradient tile (GemmDesc::ksplit: the contraction over the data rows, cut so that no TMEM accumulation chain is
// longer than ~1000 rows; ~640 on the 3xTF32 engines) are processed back to back by the SAME CTA pair: segment 0 stores the tile, every later segment
// adds into it (TMA reduce-add in L2 / read-modify-write by the thread that wrote the element), in segment order - one
// gradient copy in HBM, summed in a fixed order.
struct Item { const GemmDesc* g; int chain, tile_m, tile_n, split, kb0, nkb, nsplit; };

// 3xFP16 split: scale exponent of an operand of `chain` (GemmDesc::aSc / bSc)
__device__ __forceinline__ int operand_exp(const DevState* S, int sc, int chain) {
  if (sc >= 0) return f16_scale_exp(S->amax[(size_t)sc * S->C_alloc + chain]);
  if (sc == SC_UNIT) return kF16TopExp;
  return S->xexp[sc == SC_XTRAIN ? 0 : 1];
}
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }   // 2^e, -126 <= e <= 127

__device__ __forceinline__ int items_per_chain(const DevState* S, const PhaseDesc& ph, int& ta) {
  const GemmDesc& ga = S->gemm[ph.g0];
  ta = ga.tm * ga.tn;
  const int tb = (ph.g1 >= 0) ? S->gemm[ph.g1].tm * S->gemm[ph.g1].tn : 0;
  return ta + tb;
}

__device__ __forceinline__ void set_split(Item& it, int sp) {
  const GemmDesc& g = *it.g;
  const int nkb_all = (g.K + kBK - 1) / kBK;
  it.split = sp;
  it.kb0 = (int)(((long long)nkb_all * sp) / g.ksplit);
  it.nkb = (int)(((long long)nkb_all * (sp + 1)) / g.ksplit) - it.kb0;
}

__device__ __forceinline__ bool decode_item(const DevState* S, const PhaseDesc& ph, int c_begin, int item, Item& it) {
  int ta;
  const int per_chain = items_per_chain(S, ph, ta);
  it.chain = c_begin + item / per_chain;
  int t = item % per_chain;
  it.g = (t < ta) ? &S->gemm[ph.g0] : &S->gemm[ph.g1];
  if (t >= ta) t -= ta;
  const GemmDesc& g = *it.g;
  it.tile_m = t / g.tn;
  it.tile_n = t % g.tn;
  it.nsplit = g.ksplit;
  set_split(it, 0);
  return !(g.needLang && !S->ch.is_lang[it.chain]);
}

__device__ __forceinline__ int phase_items(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end) {
  int ta;
  return items_per_chain(S, ph, ta) * (c_end - c_begin);
}

// Which items this CTA pair processes, in order: the host-computed longest-first schedule of the phase (bae_api.cu:
// build_schedule; weight-gradient tiles are ~4x longer than the other tiles of a backward phase) when the launch covers
// the chain count the schedule was built for, round-robin otherwise (parity probe, a short last chain group).
struct Sched { const int* list; int cnt, pair, npairs, total; };
__device__ __forceinline__ Sched make_sched(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end) {
  Sched sc;
  sc.pair = (int)(blockIdx.x >> 1);
  sc.npairs = (int)(gridDim.x >> 1);
  sc.total = phase_items(S, ph, c_begin, c_end);
  const bool tab = S->sched != nullptr && ph.sched_off >= 0 && (c_end - c_begin) == S->sched_chains && sc.npairs == S->sched_pairs;
  sc.list = tab ? S->sched + ph.sched_off + (size_t)sc.pair * ph.sched_max : nullptr;
  sc.cnt = tab ? ph.sched_max : 0;
  return sc;
}
__device__ __forceinline__ int sched_next(const Sched& sc, int k) {
  if (sc.list) return k < sc.cnt ? sc.list[k] : -1;     // lists are padded with -1
  const int item = sc.pair + k * sc.npairs;
  return item < sc.total ? item : -1;
}

// ---------------------------------------------------------------------------------------------------
// One GEMM phase on the tensor cores, front half: TMA producer, MMA issuer and the hi/lo converters
// (warpgroups 0 and 3). Returns when this CTA's last MMA has been issued / last lo tile written.
// ---------------------------------------------------------------------------------------------------
// a phase whose converter warps send converted tiles to an activation image (GemmDesc::aOut)
__device__ __forceinline__ bool phase_has_out(const DevState* S, const PhaseDesc& ph) {
  return S->gemm[ph.g0].aOut != nullptr || (ph.g1 >= 0 && S->gemm[ph.g1].aOut != nullptr);
}

// kMode: what the GEMMs of the phase do with operand images, fixed at compile time where the host knows it for every GEMM of
// the phase (bae_kernels.cu: launch_phase) - the converter loop at 48 registers is sensitive to every live value and every
// inlined copy of the conversion code (measured: the mere presence of the image bookkeeping cost the converter-only engine 10%
// per step, the presence of the store bookkeeping alone 5%):
//   FM_GENERIC  everything decided per GEMM at run time (3xTF32, the persistent kernel, images switched off, mixed phases)
//   FM_B        B operands are images, A operands are converted (backward phases; forward phases over the test matrix)
//   FM_AB       both operands are images: the converter warps only relay the barrier (first layer: data x weights)
//   FM_B_OUT    B operands are images, the converted A tiles are persisted as an activation image (forward, training matrix)
// kAK: every A operand of the phase is K-major (forward layers): no MN-major copy of the conversion code.
//   FM_NONE     no GEMM of the phase touches an image (3xTF32 engines, BAE_TC_PRECONV=0): the converter-only engine as it was
enum FrontMode : int { FM_GENERIC = 0, FM_B = 1, FM_AB = 2, FM_B_OUT = 3, FM_NONE = 4 };
#pragma nv_diag_suppress 128   // (a mode fixed at compile time leaves the other engine's loops unreachable)
template <int kMode, bool kAK>
__device__ void gemm_phase_front(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, Smem& sm, Ring& ring) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Sched sched = make_sched(S, ph, c_begin, c_end);
  const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(S->tmaps);
  // (the timeline trace is compiled into the generic kernels only; with BAE_TC_TRACE=1 the host launches those)
  long long* dbg = (kMode == FM_GENERIC && blockIdx.x == 0 && (S->dbg_phase < 0 || S->dbg_phase == (int)(&ph - S->phase))) ? S->dbg : nullptr;
  const uint32_t rank = cluster_rank();
  // NOTE: every descriptor field used inside a K loop is copied into a local first. The loops are full of
  // asm volatile statements with memory clobbers, after which the compiler would re-load fields of *it.g from
  // global memory (an L2 round trip each) on every K block.

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
    {
      int ntr = 0;
      for (int ki = 0;; ++ki) {   // items go to CTA pairs
        const int item = sched_next(sched, ki);
        if (item < 0) break;
        Item it;
        if (!decode_item(S, ph, c_begin, item, it)) continue;
        const GemmDesc& g = *it.g;
        for (int sp = 0; sp < it.nsplit; ++sp) {
        set_split(it, sp);
        // this CTA's share of the pair's 256 x BN tile: 128 rows of A, BN/2 columns of B
        const int m0 = it.tile_m * (2 * kBM) + (int)rank * kBM, n0 = it.tile_n * g.tcBN + (int)rank * (g.tcBN >> 1);
        // chain coordinate of the tensor maps: activations / deltas are indexed by the slot in the resident chain group
        const int ca = g.sA ? (g.aAct ? it.chain - c_begin : it.chain) : 0, cb = g.sB ? (g.bAct ? it.chain - c_begin : it.chain) : 0;
        const int dbgf = BAE_DBGF(S);
        // pre-split operand images: this CTA's hi tile of K block kb0 (the lo tile lies kbs / 2 further)
        const unsigned char* imgA = (kMode == FM_B || kMode == FM_B_OUT || kMode == FM_NONE) ? nullptr : g.aImg ? g.aImg + (size_t)ca * g.sAImg + (size_t)it.kb0 * g.aKbs + (size_t)(m0 >> 3) * 256 : nullptr;   // (A images are K-major: the data matrix)
        const unsigned char* imgB = kMode == FM_NONE ? nullptr : g.bImg ? g.bImg + (size_t)cb * g.sBImg + (size_t)it.kb0 * g.bKbs + (size_t)(n0 >> 3) * (g.bMode ? 128 : 256) : nullptr;
        const uint32_t aKbs = (uint32_t)g.aKbs, bKbs = (uint32_t)g.bKbs, bPlane = (uint32_t)(g.tcBN >> 1) * 32u;
        const uint32_t bytes = ((dbgf & 64) ? 0u : (uint32_t)kATileBytes) + ((dbgf & 32) ? 0u : (((kMode != FM_GENERIC && kMode != FM_NONE) || imgB) ? 2u * bPlane : (uint32_t)g.tcBBytes));
        const int aMode = g.aMode, bMode = g.bMode, nchB = g.tcBBytes / kMnChunkBytes, nkb = it.nkb, kb0 = it.kb0;
        const CUtensorMap* mapA = &maps[g.tmIdx[0]];
        const CUtensorMap* mapB = &maps[g.tmIdx[1]];
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&sm.empty()[ring.stage], ring.phase ^ 1);
          if (elect_one()) {
            trace(dbg, 0, ntr, kb);
            uint8_t* st = sm.tiles + ring.stage * kStageBytes;
            // image tiles take the operand's half of the RAW stage ([hi 4 KB][lo 4 KB]; the tensor core reads them there)
            const uint32_t cv = smem_u32(st);
            mbar_expect_tx(&sm.full()[ring.stage], bytes);
            const int k0 = (kb0 + kb) * kBK;
            if (dbgf & 64) {
            } else if (kMode == FM_AB || imgA) {   // (fast modes: the host has checked every GEMM of the phase)
              const unsigned char* p = imgA + (size_t)kb * aKbs;
              bulk_load(cv, p, (uint32_t)kF16TileBytes, &sm.full()[ring.stage]);
              bulk_load(cv + kF16TileBytes, p + (aKbs >> 1), (uint32_t)kF16TileBytes, &sm.full()[ring.stage]);
            } else if (aMode == 0) {
              tma_load_3d(st, mapA, &sm.full()[ring.stage], k0, m0, ca);
            } else {
              // MN-major: one [16 K-rows][32 floats] box per 32-wide chunk of the tile dimension
#pragma unroll
              for (int c = 0; c < kBM / 32; ++c)
                tma_load_3d(st + c * kMnChunkBytes, mapA, &sm.full()[ring.stage], m0 + c * 32, k0, ca);
            }
            uint8_t* sb = st + kATileBytes;
            if (dbgf & 32) {
            } else if ((kMode != FM_GENERIC && kMode != FM_NONE) || imgB) {
              const unsigned char* p = imgB + (size_t)kb * bKbs;
              if (bMode == 0) {
                bulk_load(cv + kATileBytes, p, bPlane, &sm.full()[ring.stage]);
                bulk_load(cv + kATileBytes + kF16TileBytes, p + (bKbs >> 1), bPlane, &sm.full()[ring.stage]);
              } else {   // MN-major image: one run per K half; the halves lie 128 T8 = kbs / 4 bytes apart
                const uint32_t run = bPlane >> 1, jstep = bKbs >> 2;
                bulk_load(cv + kATileBytes, p, run, &sm.full()[ring.stage]);
                bulk_load(cv + kATileBytes + run, p + jstep, run, &sm.full()[ring.stage]);
                bulk_load(cv + kATileBytes + kF16TileBytes, p + (bKbs >> 1), run, &sm.full()[ring.stage]);
                bulk_load(cv + kATileBytes + kF16TileBytes + run, p + (bKbs >> 1) + jstep, run, &sm.full()[ring.stage]);
              }
            } else if (bMode == 0) {
              tma_load_3d(sb, mapB, &sm.full()[ring.stage], k0, n0, cb);
            } else {
              for (int c = 0; c < nchB; ++c)
                tma_load_3d(sb + c * kMnChunkBytes, mapB, &sm.full()[ring.stage], n0 + c * 32, k0, cb);
            }
          }
          __syncwarp();
          ring.advance();
        }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of the pair; warp-uniform loop, one elected lane issues) =================
    if (rank == 0) {
      int ntr = 0;
      const int dbgf = BAE_DBGF(S);
      const uint32_t tmem = *sm.tmem_base();
      for (int ki = 0;; ++ki) {   // items go to CTA pairs
        const int item = sched_next(sched, ki);
        if (item < 0) break;
        Item it;
        if (!decode_item(S, ph, c_begin, item, it)) continue;
        const GemmDesc& g = *it.g;
        for (int sp = 0; sp < it.nsplit; ++sp) {
        set_split(it, sp);
        mbar_wait_cluster(&sm.tmem_empty()[0], ring.acc_phase ^ 1);   // the epilogue warps of BOTH CTAs have drained the previous tile
        fence_after_sync();
        const uint32_t d_main = tmem, d_cross = tmem + (uint32_t)g.tcBN;
        const bool amn = g.aMode != 0, bmn = g.bMode != 0;
        const uint32_t astep = amn ? (1024u >> 4) : (32u >> 4), bstep = bmn ? (1024u >> 4) : (32u >> 4);
        const uint32_t idesc = g.idesc;
        const int nkb = it.nkb, krem0 = g.K - it.kb0 * kBK;
        // descriptors of stage 0; a stage / K-step offset only moves the 14-bit address field (bytes >> 4)
        const uint64_t dA = make_desc(smem_u32(sm.tiles), amn), dB = make_desc(smem_u32(sm.tiles) + kATileBytes, bmn);
        const uint64_t d16 = make_desc_f16(smem_u32(sm.tiles));
        const bool f16 = (kMode != FM_GENERIC && kMode != FM_NONE) || S->f16 != 0;   // (operand images exist on the 3xFP16 engine only)
        const bool preA = kMode == FM_GENERIC ? g.aImg != nullptr : kMode == FM_AB, preB = kMode == FM_GENERIC ? g.bImg != nullptr : kMode != FM_NONE;   // operand images: read in the raw stage, which these MMAs then release
        // B image in the raw stage: K-major [t8][j][r] like a converted tile; MN-major [j][t8][r] (strides: K half 16 n, MN group 128)
        const uint64_t rawB = bmn ? make_desc_f16(smem_u32(sm.tiles) + kATileBytes, (uint32_t)(g.tcBN >> 1) * 16u, 128u)
                                  : make_desc_f16(smem_u32(sm.tiles) + kATileBytes);
        for (int kb = 0; kb < nkb; ++kb) {
          // `conv` is signalled by converter warps that had themselves waited for the TMA bytes (`full`)
          mbar_wait(&sm.conv()[ring.lo], ring.lo_phase);   // both CTAs' raw and lo tiles of the K block are ready
          fence_after_sync();
          if (f16) {
            // 3xFP16: the converted stage holds [A hi][A lo][B hi][B lo], 4 KB each; one K = 16 MMA per product
            if (elect_one()) {
              trace(dbg, 1, ntr, kb);
              const uint64_t cvt = d16 + (kStages + ring.lo) * (uint32_t)(kStageBytes >> 4);
              const uint64_t ah = preA ? d16 + ring.stage * (uint32_t)(kStageBytes >> 4) : cvt, al = ah + (kF16TileBytes >> 4);
              const uint64_t bh = preB ? rawB + ring.stage * (uint32_t)(kStageBytes >> 4) : cvt + (2 * kF16TileBytes >> 4), bl = bh + (kF16TileBytes >> 4);
              const uint32_t acc0 = kb ? 1u : 0u;
              umma_f16(d_main, ah, bh, idesc, acc0);
              umma_f16(d_cross, al, bh, idesc, acc0);
              umma_f16(d_cross, ah, bl, idesc, 1u);
              umma_commit(&sm.lo_empty()[ring.lo]);
              if (preA || preB) umma_commit(&sm.empty()[ring.stage]);   // (both CTAs' stages, like every commit here)
              if (kb == nkb - 1) {
                trace(dbg, 1, ntr, 1000);
                umma_commit(&sm.tmem_full()[0]);
              }
            }
          } else
          if (elect_one()) {
            trace(dbg, 1, ntr, kb);
            const uint32_t so = ring.stage * (uint32_t)(kStageBytes >> 4), lo = (kStages + ring.lo) * (uint32_t)(kStageBytes >> 4);
            const uint64_t ah = dA + so, al = dA + lo, bh = dB + so, bl = dB + lo;
            const bool two = (krem0 - kb * kBK) > 8;   // second 8-wide K step present (not beyond K)
            const uint32_t acc0 = kb ? 1u : 0u;
            umma_tf32(d_main, ah, bh, idesc, acc0);
            if (two) umma_tf32(d_main, ah + astep, bh + bstep, idesc, 1u);
            if (!(dbgf & 8)) {
              umma_tf32(d_cross, al, bh, idesc, acc0);
              umma_tf32(d_cross, ah, bl, idesc, 1u);
              if (two) {
                umma_tf32(d_cross, al + astep, bh + bstep, idesc, 1u);
                umma_tf32(d_cross, ah + astep, bl + bstep, idesc, 1u);
              }
            }
            umma_commit(&sm.empty()[ring.stage]);    // frees the raw stage (in both CTAs) when these MMAs have read it
            umma_commit(&sm.lo_empty()[ring.lo]);    // ... and the lo stage
            if (kb == nkb - 1) {
              trace(dbg, 1, ntr, 1000);
              umma_commit(&sm.tmem_full()[0]);      // accumulators complete -> epilogue
            }
          }
          __syncwarp();
          ring.advance();
        }
        ring.acc_phase ^= 1;
        }
      }
    }
  } else {
    // ================= hi/lo converters: groups {2,3}, {12,13}, {14,15}; K block k goes to group k % 3 =================
    // lo = x - trunc_tf32(x), element-wise on the swizzled tile bytes (position independent). A raw pair is
    // [A][B] contiguous and the lo pair has the same layout, so one flat loop covers both operands; the pair is a
    // multiple of 64 float4, so the 64 threads of a group run without predicates.
    const int grp = (warp < 4) ? 0 : 1 + ((warp - 12) >> 1);
    const int gt = ((warp & 1) << 5) + lane;   // 0..63 inside the group
    const uint32_t tiles_s = smem_u32(sm.tiles) + gt * 16;
    const bool skip = (BAE_DBGF(S) & 1) != 0;
    const bool f16 = (kMode != FM_GENERIC && kMode != FM_NONE) || S->f16 != 0;
    const uint32_t tiles0 = smem_u32(sm.tiles);
    const ConvLane cl = conv_lane(gt);
    int ntr = (grp == 0) ? 0 : (grp == 1 ? 170 : 340);
    const bool preMode = kMode == FM_GENERIC ? S->preconv != 0 : kMode != FM_NONE;   // barrier counts of this launch (setup)
    const bool outMode = kMode == FM_B_OUT || (kMode == FM_GENERIC && phase_has_out(S, ph));
    uint32_t kbc = ring.kbc;
    for (int ki = 0;; ++ki) {   // items go to CTA pairs
      const int item = sched_next(sched, ki);
      if (item < 0) break;
      Item it;
      if (!decode_item(S, ph, c_begin, item, it)) continue;
      const int nkb = (it.g->K + kBK - 1) / kBK;   // all K segments of the tile, back to back
      // 3xFP16: operand scales of this item, source majors, units of this CTA's B half (8 columns x 16 k each)
      float sA = 1.f, sB = 1.f;
      uint32_t offA = 0, offB = 0;
      bool swapA = false, swapB = false, aM = false, bM = false;
      int nvB = 0;   // units of this CTA's B half that this thread converts: gt + 64 n < tcBN
      const bool preA = kMode == FM_GENERIC ? it.g->aImg != nullptr : kMode == FM_AB, preB = kMode == FM_GENERIC ? it.g->bImg != nullptr : kMode != FM_NONE;   // the producer delivers these operands already split
      // A operand persisted as an activation image (GemmDesc::aOut): after every K block this group sends the converted tiles
      // [A hi], [A lo] to the image (one box store each); the slot may be rewritten once the stores have READ it - the second
      // arrival on `lo_empty` (next to the MMAs' commit), made when this group comes back for its next block
      // (the hi plane by the tile_n == 0 item of a row block, the lo plane by its tile_n == 1 item when there is one: the items
      //  of a phase then cost the same, which its schedule assumes)
      const bool outK = kMode == FM_B_OUT || (kMode == FM_GENERIC && f16 && it.g->aOut);
      const bool outHi = outK && it.tile_n == 0, outLo = outK && it.tile_n == (it.g->tn > 1 ? 1 : 0);
      const CUtensorMap* mapOut = (outHi || outLo) ? &maps[it.g->tmOut] : nullptr;
      const int outRow = (it.tile_m * (2 * kBM) + (int)rank * kBM) >> 4, outSlot = it.chain - c_begin;
      if (f16) {
        sA = exp2i(operand_exp(S, it.g->aSc, it.chain));
        sB = exp2i(operand_exp(S, it.g->bSc, it.chain));
        aM = !kAK && it.g->aMode != 0; bM = it.g->bMode != 0;
        offA = cl.off[aM ? 1 : 0]; offB = cl.off[bM ? 1 : 0];
        swapA = aM && cl.swap1; swapB = bM && cl.swap1;
        nvB = (it.g->tcBN - gt + 63) >> 6;
      }
      if (f16) {
        // This group's K blocks of the item, without touching the others: block kb of the item is block bk + kb of the CTA's
        // stream, which fixes its group ((bk + kb) % 3), its raw stage and its converted slot. (A per-block loop that merely
        // skipped the other groups' blocks cost ~250 cycles per skipped block: its loop-carried state lives in local memory.)
        for (int kb = (grp + kConvGroups - (int)kbc) % kConvGroups; kb < nkb; kb += kConvGroups) {
          const uint32_t b7 = ring.stage + (uint32_t)kb, b5 = ring.lo + (uint32_t)kb;
          const uint32_t stage = b7 % kStages, phase = ring.phase ^ ((b7 / kStages) & 1u);
          const uint32_t lo = b5 % kLoStages, lo_phase = ring.lo_phase ^ ((b5 / kLoStages) & 1u);
          mbar_wait(&sm.lo_empty()[lo], lo_phase ^ 1);
          mbar_wait(&sm.full()[stage], phase);
          if (gt == 0) trace(dbg, 2, ntr, kb);
          const uint32_t src = tiles0 + stage * kStageBytes, dst = tiles0 + (kStages + lo) * kStageBytes + gt * 16;
          // A: 256 units = 4 per thread; B: tcBN units (a multiple of 16) = this CTA's BN/2 columns, nvB of them this thread's
          if ((BAE_DBGF(S) & 3) || preA) {
          } else if (aM) conv_tile<true>(src + offA, dst, dst + kF16TileBytes, swapA, sA, 4);
          else conv_tile<false>(src + offA, dst, dst + kF16TileBytes, false, sA, 4);
          if ((BAE_DBGF(S) & 5) || preB) {
          } else if (bM) conv_tile<true>(src + kATileBytes + offB, dst + 2 * kF16TileBytes, dst + 3 * kF16TileBytes, swapB, sB, nvB);
          else conv_tile<false>(src + kATileBytes + offB, dst + 2 * kF16TileBytes, dst + 3 * kF16TileBytes, false, sB, nvB);
          if (gt == 0) trace(dbg, 2, ntr, 2000 + kb);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (mapOut && !(BAE_DBGF(S) & 256)) asm volatile("bar.sync %0, 64;" ::"r"(3 + grp) : "memory");   // both warps of the group have written (and fenced) the tile
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&sm.empty()[stage]);   // the raw stage has been read: the producer may refill it
            // (third arrival: the MMAs' commit when the stage holds an image tile, else this group's first warp again)
            if (preMode && !(preA || preB) && (warp & 1) == 0) mbar_arrive(&sm.empty()[stage]);
            if (rank == 0) mbar_arrive(&sm.conv()[lo]);
            else mbar_arrive_leader_relaxed(&sm.conv()[lo]);
            if (outMode && (warp & 1) == 0) {   // gt == 0
              if (mapOut && !(BAE_DBGF(S) & 128)) {
                const uint32_t cvs = tiles0 + (kStages + lo) * kStageBytes;
                if (outHi) tma_store_4d(mapOut, cvs, (it.kb0 + kb) * 128, 0, outRow, outSlot);
                if (outLo) tma_store_4d(mapOut + 1, cvs + kF16TileBytes, (it.kb0 + kb) * 128, 0, outRow, outSlot);
                bulk_commit();
                // (waiting here costs this thread the few hundred cycles the copy engine needs to read 4 KB; leaving the arrival to
                //  the group's next block - three blocks later - kept three of the five converted slots occupied)
                bulk_wait_read();
                mbar_arrive(&sm.lo_empty()[lo]);
              } else {
                mbar_arrive(&sm.lo_empty()[lo]);   // nothing but the MMAs reads the slot
              }
            }
          }
        }
        // the item's blocks are done (or left to the other groups): move the ring state past them
        const uint32_t e7 = ring.stage + (uint32_t)nkb, e5 = ring.lo + (uint32_t)nkb;
        ring.stage = e7 % kStages; ring.phase ^= (e7 / kStages) & 1u;
        ring.lo = e5 % kLoStages; ring.lo_phase ^= (e5 / kLoStages) & 1u;
        kbc = (kbc + (uint32_t)nkb) % kConvGroups;
        continue;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        if ((int)kbc == grp) {
          // The lo slot first (the MMAs of K block k - 4 are done): it also pins the phase of `full`. Parity waits cannot
          // tell phases two apart, and a raw stage is served by all groups in turn; the MMAs of block k - 4 complete after
          // those of k - 7, which consumed this raw stage's previous contents.
          if (gt == 0) trace(dbg, 2, ntr, 7000 + kb);
          mbar_wait(&sm.lo_empty()[ring.lo], ring.lo_phase ^ 1);
          if (gt == 0) trace(dbg, 2, ntr, 8000 + kb);
          mbar_wait(&sm.full()[ring.stage], ring.phase);
          if (gt == 0) trace(dbg, 2, ntr, kb);
          const uint32_t src = tiles_s + ring.stage * kStageBytes;
          const uint32_t dst = tiles_s + (kStages + ring.lo) * kStageBytes;
          if (!skip) {
            // the whole 16 KB tile pair, 16 float4 per thread in two batches of 8 (bytes past a narrow B half are
            // converted too: nobody reads them)
#pragma unroll
            for (int i = 0; i < kStageBytes / (16 * kConvGroupThreads); i += 8) {
              float4 v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = lds128(src + (i + u) * (16 * kConvGroupThreads));
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                float h;
                split_tf32(v[u].x, h, v[u].x); split_tf32(v[u].y, h, v[u].y);
                split_tf32(v[u].z, h, v[u].z); split_tf32(v[u].w, h, v[u].w);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) sts128(dst + (i + u) * (16 * kConvGroupThreads), v[u]);
            }
          }
          if (gt == 0) trace(dbg, 2, ntr, 2000 + kb);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> tensor-core (async proxy) reads of THIS CTA's shared memory
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&sm.conv()[ring.lo]);
            else mbar_arrive_leader_relaxed(&sm.conv()[ring.lo]);
          }
          if (gt == 0) trace(dbg, 2, ntr, 3000 + kb);
        }
        if (++kbc == (uint32_t)kConvGroups) kbc = 0;
        ring.advance();
      }
    }
    ring.kbc = kbc;
  }
}

// ---------------------------------------------------------------------------------------------------
// One GEMM phase, back half: the eight epilogue warps (warpgroups 1 and 2). Returns after every tile of
// this CTA is stored.
// ---------------------------------------------------------------------------------------------------
struct EpiCtx {
  int q, eg, ew, et, lane, sub_r, sub_c;
  uint32_t tmem, stg_s, rank;
  long long* dbg;
  int dbgf;
};

// One tile, specialised by the epilogue kinds of the phase: the body holds only their code around the 128 accumulator
// registers (a body with all six kinds selected at run time spilled inside the store loop, an L2 round trip per spill).
template <int kEpi, int kEpiB>
__device__ __forceinline__ void epilogue_tile(const DevState* S, const Item& it, const EpiCtx& c, Smem& sm, Ring& ring, int& ntr, int c_begin) {
  const GemmDesc& g = *it.g;
  const int tcBN = g.tcBN;
  const int n0 = it.tile_n * tcBN;
  const int n_end = min(g.N, n0 + tcBN);   // this tile owns output columns [n0, n_end)
  const int ncols = n_end - n0;
  const int eg = c.eg, lane = c.lane, sub_r = c.sub_r, sub_c = c.sub_c;
  const uint32_t tbase = c.tmem + ((uint32_t)(c.q * 32) << 16);
  long long* dbg = c.dbg;
  // kEpi < 0: kind selected at run time (persistent kernel, debug); one kind: every test on `epi` below folds away; two
  // kinds (a backward phase: weight gradient | delta): ONE body that keeps only these two kinds' code - two inlined bodies
  // made the compiler keep one of the accumulator arrays in local memory
  const int epi = (kEpi < 0) ? g.epi : (kEpiB < 0 ? kEpi : (g.epi == kEpi ? kEpi : kEpiB));
  const bool kHasBias = (epi == EPI_SIGMOID || epi == EPI_BIAS || epi == EPI_LOSS);
  const bool kHasAux = (epi == EPI_DSIG || epi == EPI_LOSS);

  // bias of this thread's four output columns in each of its four chunks: requested before the accumulators are
  // waited for (an L2 round trip per chunk otherwise, in a chunk loop that is not unrolled)
  float4 b4s[4];
  if (kHasBias) {
    const float* bias0 = g.bias + (size_t)it.chain * g.sBias;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = n0 + (eg + 2 * j) * 32 + sub_c;
      b4s[j] = (cc < n_end) ? __ldg(reinterpret_cast<const float4*>(&bias0[cc])) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // 3xFP16: undo the operand scales (powers of two). 3xTF32: 1, 1.
  float unscale0 = 1.f, unscale1 = 1.f;
  if (S->f16) {
    const int e = operand_exp(S, g.aSc, it.chain) + operand_exp(S, g.bSc, it.chain);
    unscale0 = unscale1 = exp2i(-e);
  }

  // ---- drain: both accumulators of this warp's chunks -> registers, then TMEM goes back to the MMA issuer.
  mbar_wait(&sm.tmem_full()[0], ring.acc_phase);
  if (c.et == 0) trace(dbg, 3, ntr, 0);
  fence_after_sync();
  uint32_t acc[4][32];
  {
    // all main-accumulator chunks are requested at once, then the cross-term chunks 16 columns at a time.
    // (Handing the main accumulator back first, so that the next tile's hi*hi products start during the cross-term drain,
    // was measured twice - on the converter-bound engine: the gap between tiles shrank by 1.7 k cycles, the step did not get
    // faster; on the operand-image engine, with the hi*hi products of the first four K blocks issued ahead of the cross
    // terms: again no change of the step time - not kept.)
    uint32_t pa[16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((eg + 2 * j) * 32 < ncols) tmem_ld32_issue(tbase + (uint32_t)((eg + 2 * j) * 32), acc[j]);
    // (the first waiting load below also completes the main-accumulator loads)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = (eg + 2 * j) * 32;
      if (c0 < ncols) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld16_issue(tbase + (uint32_t)(tcBN + c0 + hh * 16), pa);
          tmem_wait_ld16(pa);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            // (acc[j] is the target of an asynchronous tcgen05.ld: every instruction that reads it must also depend on `pa`,
            //  whose wait completes all loads issued so far - hence the product with pa first, then ONE fma that reads acc)
            acc[j][hh * 16 + i] = __float_as_uint(fmaf(__uint_as_float(acc[j][hh * 16 + i]), unscale0, __uint_as_float(pa[i]) * unscale1));
        }
      }
    }
  }
  fence_before_sync();
  __syncwarp();
  // The leader's MMA issuer waits for both CTAs' epilogue warps. What is handed over is tensor memory whose loads have
  // COMPLETED (tcgen05.wait::ld above): no shared / global memory of this warp has to become visible to the issuer, so the
  // arrival is a plain signal (a release at cluster scope costs ~1500 cycles per arrival, measured on the converters;
  // GEMM phases -3.0%). Handing the main accumulator back before the cross-term drain was measured again on top of
  // this (two hand-back barriers): no further gain.
  if (lane == 0) {
    if (c.rank == 0) mbar_arrive(&sm.tmem_empty()[0]);
    else mbar_arrive_leader_relaxed(&sm.tmem_empty()[0]);
  }
  ring.acc_phase ^= 1;
  if (c.et == 0) trace(dbg, 3, ntr, 1);

  // locals: see the note in gemm_phase_front
  const int chain = it.chain;
  const int m0 = it.tile_m * (2 * kBM) + (int)c.rank * kBM;   // this CTA's 128 rows of the pair's tile
  // K segments after the first ADD into what this CTA pair stored for the previous ones (same warp, same lane per element)
  const bool accum = (epi == EPI_GRADW) && it.split > 0;
  const int slot = chain - c_begin;        // the chain's slot in the resident group (activation / delta buffers)
  float* C = g.C + (size_t)(g.cAct ? slot : chain) * g.sC;   // direct stores: only chunks that straddle the next tile's columns
  const int ldc = g.ldc;
  const CUtensorMap* mapC = reinterpret_cast<const CUtensorMap*>(S->tmaps) + g.tmIdx[2];
  const int zC = g.sC ? (g.cAct ? slot : chain) : 0;   // slice of the output tensor map
  const float* aux = (kHasAux && g.aux) ? g.aux + (size_t)(g.auxAct ? slot : chain) * g.sAux : nullptr;
  const int gM = g.M, ldaux = g.ldaux;
  const int biasCol = (epi == EPI_GRADW) ? g.biasCol : -1;
  float* gbias = (epi == EPI_GRADW && g.gbias) ? g.gbias + (size_t)chain * g.sGb : nullptr;
  const float scale = g.scale;
  const int r0 = m0 + c.q * 32;
  const uint32_t stg_s = c.stg_s;

  // ---- fused epilogue from registers, one 32 x 32 chunk at a time. The chunk loop is NOT unrolled (unrolled, this
  // section was 160 KB of straight-line code and stalled on instruction fetch) and nothing may spill. The chunk's
  // accumulators go to the staging tile, and all epilogue math runs in the coalesced (transposed) domain where aux
  // loads, bias and stores share one address pattern.
  float l_sse = 0.f, l_sd = 0.f;
  float amx = 0.f;   // max |stored value| of this thread (deltas: bound behind the FP16 scale of the next GEMMs' operand)
  const bool kTrackAmax = (epi == EPI_LOSS || epi == EPI_DSIG || epi == EPI_PLAIN);
  const int lim = (epi == EPI_GRADW && biasCol >= 0) ? min(n_end, biasCol) : n_end;   // columns >= lim are not stored
  if (accum) {
    // the previous segment's box stores of this warp are COMPLETE (not just read) before anything is added to them; they
    // were issued a whole main loop ago, so this does not wait in practice
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }
#pragma unroll 1
  for (int jj = 0; jj < 4; ++jj) {
    const int c0 = (eg + 2 * jj) * 32;
    if (c0 >= ncols || (c.dbgf & 16)) break;
    const int cc = n0 + c0 + sub_c;            // first of this thread's 4 output columns
    const bool cv = cc < n_end;
    // aux tile (sigmoid outputs for EPI_DSIG, data for EPI_LOSS), 4 rows x 128 B per instruction, in two groups of four
    float4 t[4];
    const float* ap = kHasAux ? aux + (size_t)(r0 + sub_r) * ldaux + cc : nullptr;
    auto load_aux = [&](int half) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ii = half * 4 + i;
        t[i] = (cv && r0 + ii * 4 + sub_r < gM) ? ldg128_pinned(ap + (size_t)(ii * 4) * ldaux) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (kHasAux) load_aux(0);
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kHasBias) {
      b4 = b4s[0];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (j == jj) b4 = b4s[j];
    }
    if (c.et == 0) trace(dbg, 3, ntr, 10 + jj);
    // only chunks whose 32 columns are all stored go through the copy engine; the last, partial chunk of a tile (tile widths
    // are multiples of 16; the tensor's edge; the bias-gradient column) takes the masked direct-store path. The copy engine
    // does clip at the tensor's edge, but in 16-byte units: with a width of 450 it wrote columns 450 and 451 as well.
    const bool box = n0 + c0 + 32 <= lim;
    // the previous chunk's box store must have read the staging tile before it is rewritten
    if (lane == 0) bulk_wait_read();
    __syncwarp();
    // Kinds without an aux tile need no pass in the coalesced domain when the chunk leaves as a box: bias (a broadcast load per
    // four columns) and sigmoid are applied on the way INTO the staging tile, which then already is the box to store - one
    // shared-memory pass instead of three (shared memory is the unit the main loop is short of).
    const bool direct = box && (epi == EPI_GRADW || epi == EPI_PLAIN || epi == EPI_SIGMOID || epi == EPI_BIAS);
    const float* biasc = (kHasBias && direct) ? g.bias + (size_t)chain * g.sBias + n0 + c0 : nullptr;
    // dense 32 x 128 B tile, 16-byte units XOR-swizzled by the row (the layout the store's tensor map expects)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j == jj) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 o = make_float4(__uint_as_float(acc[j][j4 * 4]), __uint_as_float(acc[j][j4 * 4 + 1]),
                                 __uint_as_float(acc[j][j4 * 4 + 2]), __uint_as_float(acc[j][j4 * 4 + 3]));
          if (biasc) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(biasc + j4 * 4));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            if (epi == EPI_SIGMOID) {
              o.x = __fdividef(1.0f, 1.0f + __expf(-o.x)); o.y = __fdividef(1.0f, 1.0f + __expf(-o.y));
              o.z = __fdividef(1.0f, 1.0f + __expf(-o.z)); o.w = __fdividef(1.0f, 1.0f + __expf(-o.w));
            }
          }
          // (rows >= M of the tile hold zero accumulators: they cannot raise the bound)
          if (direct && epi == EPI_PLAIN) amx = fmaxf(amx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
          sts128(stg_s + lane * 128 + ((j4 ^ (lane & 7)) << 4), o);
        }
      }
    __syncwarp();
    if (c.et == 0) trace(dbg, 3, ntr, 20 + jj);
    if (!direct) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kHasAux && i == 4) load_aux(1);
      const int rr = r0 + i * 4 + sub_r;
      const uint32_t sa = stg_s + (i * 4 + sub_r) * 128 + (((lane & 7) ^ ((i * 4 + sub_r) & 7)) << 4);
      float4 o = lds128(sa);
      if (rr >= gM || !cv) continue;
      if (kHasBias) { o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w; }
      if (epi == EPI_SIGMOID) {
        o.x = __fdividef(1.0f, 1.0f + __expf(-o.x)); o.y = __fdividef(1.0f, 1.0f + __expf(-o.y));
        o.z = __fdividef(1.0f, 1.0f + __expf(-o.z)); o.w = __fdividef(1.0f, 1.0f + __expf(-o.w));
      } else if (epi == EPI_DSIG) {
        o.x *= t[i & 3].x * (1.0f - t[i & 3].x); o.y *= t[i & 3].y * (1.0f - t[i & 3].y);
        o.z *= t[i & 3].z * (1.0f - t[i & 3].z); o.w *= t[i & 3].w * (1.0f - t[i & 3].w);
      } else if (epi == EPI_LOSS) {
        float d[4] = {o.x - t[i & 3].x, o.y - t[i & 3].y, o.z - t[i & 3].z, o.w - t[i & 3].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (cc + e < n_end) { l_sse = fmaf(d[e], d[e], l_sse); l_sd += d[e]; }
          d[e] *= scale;
        }
        o = make_float4(d[0], d[1], d[2], d[3]);
      }
      // (columns past n_end of a float4 hold at most values of the same magnitude - zero accumulators, or the data's ones
      //  column under the loss epilogue: a bound, not a maximum, is needed)
      if (kTrackAmax) amx = fmaxf(amx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
      if (box) {
        sts128(sa, o);   // back to where this thread read it from
      } else {
        float* dst = &C[(size_t)rr * ldc + cc];
        if (cc + 4 <= lim) {
          if (accum) {
            const float4 p4 = *reinterpret_cast<const float4*>(dst);
            o.x += p4.x; o.y += p4.y; o.z += p4.z; o.w += p4.w;
          }
          *reinterpret_cast<float4*>(dst) = o;
        } else {
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (cc + e < lim) dst[e] = accum ? dst[e] + ov[e] : ov[e];
            else if (epi == EPI_GRADW && cc + e == biasCol) gbias[rr] = accum ? gbias[rr] + ov[e] : ov[e];
          }
        }
      }
    }
    }   // !direct
    // ONE box store per chunk through the copy engine (rows >= M are clipped by it): global
    // stores issued by the epilogue warps themselves slowed the converters' shared-memory stream down (traced: 2.6 k cycles per tile)
    if (box) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (accum) tma_reduce_add_3d(mapC, stg_s, n0 + c0, r0, zC);
        else tma_store_3d(mapC, stg_s, n0 + c0, r0, zC);
        bulk_commit();
      }
    }
    __syncwarp();
  }
  if (c.et == 0) trace(dbg, 3, ntr, 2);
  if (kTrackAmax && g.cSc >= 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
    if (lane == 0 && amx > 0.f) atomicMax(reinterpret_cast<unsigned*>(S->amax) + (size_t)g.cSc * S->C_alloc + chain, __float_as_uint(amx));
  }
  if (epi == EPI_LOSS) {
    double s0 = (double)l_sse, s1 = (double)l_sd;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (lane == 0) { sm.red()[c.ew * 2] = s0; sm.red()[c.ew * 2 + 1] = s1; }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (c.et == 0) {
      double* p = g.part + ((size_t)chain * g.partStride + (size_t)(it.tile_m * 2 + (int)c.rank) * g.tn + it.tile_n) * 2;
      double a0 = 0.0, a1 = 0.0;
      for (int w = 0; w < kEpiWarps; ++w) { a0 += sm.red()[w * 2]; a1 += sm.red()[w * 2 + 1]; }
      p[0] = a0;
      p[1] = a1;
    }
  }
}

// kEpiA / kEpiB: epilogue kinds of the phase's GEMMs (kEpiB < 0: one kind; kEpiA < 0: selected at run time)
template <int kEpiA, int kEpiB>
__device__ __forceinline__ void gemm_phase_epilogue(const DevState* S, const PhaseDesc& ph, int c_begin, int c_end, Smem& sm, Ring& ring) {
  const int warp = threadIdx.x >> 5;
  const Sched sched = make_sched(S, ph, c_begin, c_end);
  EpiCtx c;
  c.lane = threadIdx.x & 31;
  c.q = warp & 3;               // TMEM lane quarter this warp may access
  c.eg = (warp - 4) >> 2;       // 0/1: which alternating 32-column chunks this warp owns
  c.ew = warp - 4;              // 0..7
  c.et = threadIdx.x - 128;     // 0..255
  c.tmem = *sm.tmem_base();
  c.stg_s = smem_u32(sm.epi_stage()) + c.ew * 4096;   // this warp's 32 x 32 staging tile (dense, 1024-byte aligned, swizzled)
  c.sub_r = c.lane >> 3; c.sub_c = (c.lane & 7) * 4;  // coalesced view: 4 rows x 128 B per instruction
  c.dbg = (blockIdx.x == 0 && (S->dbg_phase < 0 || S->dbg_phase == (int)(&ph - S->phase))) ? S->dbg : nullptr;
  c.dbgf = BAE_DBGF(S);
  c.rank = cluster_rank();
  int ntr = 0;
  for (int ki = 0;; ++ki) {   // items go to CTA pairs
    const int item = sched_next(sched, ki);
    if (item < 0) break;
    Item it;
    if (!decode_item(S, ph, c_begin, item, it)) continue;
    for (int sp = 0; sp < it.nsplit; ++sp) {
      set_split(it, sp);
      epilogue_tile<kEpiA, kEpiB>(S, it, c, sm, ring, ntr, c_begin);
    }
  }
  if (c.lane == 0) bulk_wait_all();   // this warp's box stores are complete before the kernel (or the phase) ends
  __syncwarp();
}

}  // namespace tc
}  // namespace bae
