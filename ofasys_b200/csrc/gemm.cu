// bf16 GEMM on the 5th-generation tensor cores (tcgen05) of sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ residual[M,N]),  fp32 accumulation in TMEM.
//
// Persistent, warp-specialised kernel (one CTA per SM):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A and B into a SWIZZLE_128B smem ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16,
//               accumulators double-buffered in tensor memory so the epilogue of tile i overlaps the
//               main loop of tile i+1; owns tcgen05.alloc / dealloc
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / residual -> global
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and tmem full/empty mbarriers (MMA <-> epilogue).
//
// Operand layouts: each operand may be "K-major" (rows of K contiguous: activations [M,K], weights
// [N,K]) or "MN-major" (the transposed matrix read in place: dY^T, X^T for wgrad; W for dgrad), so
// dgrad and wgrad need no transposed copies in HBM.  Both use 128-byte swizzled smem tiles; only the
// TMA box shape and the UMMA smem/instruction descriptors differ.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kSmemMax = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 2-CTA (cta_group::2) variants -----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are signalled on the LEADER's barrier
// (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
// same, delivered to the same smem offset in every CTA of `cta_mask` (cluster ranks); each destination signals the
// barrier of ITS pair's leader.  Used by 4-CTA clusters: two CTA pairs that work on adjacent N tiles of one M tile
// share the A operand -- each CTA fetches half of its A rows' K range and multicasts it to its twin in the other pair.
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(mbar), "h"(cta_mask), "r"(c0), "r"(c1)
      : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B (layout_type 2 at bits [61,64)), version 1 at [46,48):
//   bits [0,14) address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4.
//  K-major : rows of 128 B; 8-row groups are 1024 B apart (SBO); one swizzle atom along K (LBO unused).
//  MN-major: 64-element (128 B) MN atoms of BK rows each; SBO = 8 K-rows (1024 B),
//            LBO = stride between MN atoms (BK * 128 B).
// (built in the MMA issuer: constant high word, low word advanced per K step)

struct GemmParams {
  int M, N, K;
  const bf16* bias;       // [N] or null
  const float* residual;  // [M, ldr] or null
  int64_t ldr;
  void* D;
  int64_t ldd;
  int d_bf16;
  int tma_store;  // epilogue writes bf16 tiles through swizzled smem + TMA (coalesced); else direct st.global
  // split-K: the K loop is cut into `splits` ranges of kb_per_split k-blocks; range s of output tile t is its own
  // work item and writes a partial result into slab s of D (slabs are split_stride elements apart; fp32, direct path)
  int splits, kb_per_split;
  int64_t split_stride;
  int dbg;        // development: 1 = no global stores, 2 = no TMEM reads, 4 = no MMA issue, 8 = no TMA loads
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1)
               : "memory");
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) owns a 256 x BN tile;
// each CTA stages its 128 rows of A and HALF of the B tile, the leader issues UMMA 256 x BN x 16 that reads
// both CTAs' shared memory, so per-SM smem traffic per MMA is halved.
// BK = K elements per pipeline stage (64 or 128): one mbarrier hand-shake per BK/16 MMAs; with BK = 128 a
// K-major operand stage holds two 64-wide swizzle atoms.
// CL = cluster size: CG for the plain forms; 4 = two CTA pairs per cluster (CG = 2) on N-adjacent tiles of the same
// M tile with the A operand multicast between them: A traffic from L2 is halved (the kernel is fed at ~70 B/ns per SM
// whatever the tile shape, so bytes per FLOP, not MMA issue, bound it).
template <int BN, bool A_MN, bool B_MN, int CG, int BK, int STAGES, int CL>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_d, const GemmParams p) {
  constexpr int BNL = BN / CG;  // B rows staged by this CTA
  constexpr int KA = BK / 64;   // 64-wide K atoms per stage
  constexpr uint32_t A_BYTES = BLOCK_M * BK * 2;
  constexpr uint32_t B_BYTES = BNL * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // double-buffered accumulator (256 or 512 columns)
  constexpr uint32_t STG_BYTES = BLOCK_M * 128;  // one 128-row x 64-col bf16 store tile
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "tmem columns must be a power of two");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; do not rely on the dynamic-smem base
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_st = smem_b + STAGES * B_BYTES;  // 2 store staging tiles
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_st + 2 * STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);  // tiles of 128*CG rows
  const int n_tiles = (p.N + BN - 1) / BN;
  constexpr int NP = CL / CG;              // CTA pairs (CG = 2) per cluster: 1, or 2 with A multicast
  static_assert(CL == CG || (CG == 2 && CL == 4 && BK == 128), "cluster forms: CG, or 4 = two pairs");
  const int mn_tiles = m_tiles * (n_tiles / NP);  // work items: NP N-adjacent tiles each (host guarantees n_tiles % NP == 0)
  const int num_tiles = mn_tiles * p.splits;
  const int num_k_blocks = (p.K + BK - 1) / BK;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;  // rank in the cluster
  const uint32_t rank = crank & 1u;                          // rank in the CTA pair
  const uint32_t pair = crank >> 1;                          // which pair of the cluster
  const uint32_t pair_leader = crank & ~1u;                  // cluster rank of this pair's leader CTA
  const bool leader = rank == 0;
  const int tile0 = (int)(blockIdx.x / CL);
  const int tile_step = (int)(gridDim.x / CL);
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));   // both CTAs of this pair
  const uint16_t all_mask = (uint16_t)((1u << CL) - 1u);     // every CTA of the cluster

  pdl_launch();  // PDL: the next kernel's CTAs may become resident as ours retire
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_d) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, 1);  // leader's barrier: ONE arrival (its own expect_tx, which counts the bytes of both CTAs of the pair)
      mbar_init(empty_bar + i, NP);  // one commit per pair whose MMAs read this CTA's stage (own pair; + the twin pair's A multicast target)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 4 * CG);  // one arrival per epilogue warp (of both CTAs, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG == 2) cluster_sync_all();  // both CTAs are resident before the paired TMEM allocation
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // PDL: everything above (descriptor prefetch, barrier init, cluster handshake, TMEM allocation) overlaps the tail
  // of the previous kernel; global memory (TMA loads, bias / residual reads, stores) is touched only below.
  pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer
    // (whole warp, warp-uniform control flow; one elected lane issues -- see the MMA issuer)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int tt = t % mn_tiles, sp = t / mn_tiles;
        const int m0 = (tt % m_tiles) * (BLOCK_M * CG) + (int)rank * BLOCK_M;  // this CTA's 128 rows
        const int n0 = ((tt / m_tiles) * NP + (int)pair) * BN + (int)rank * BNL;  // this CTA's slice of its pair's B tile
        const int kb_begin = sp * p.kb_per_split, kb_end = min(num_k_blocks, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          // the non-leader's loads signal the leader's barrier directly (complete_tx); bytes that land before the
          // leader's expect_tx only drive the transaction count negative for a moment
          if (leader && issuer) mbar_expect_tx(full_bar + stage, (p.dbg & 8) ? 0u : (A_BYTES + B_BYTES) * CG);
          uint8_t* sa = smem_a + stage * A_BYTES;
          uint8_t* sb = smem_b + stage * B_BYTES;
          const int k0 = kb * BK;
          auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1) {
            if ((p.dbg & 8) || !issuer) return;
            if (CG == 1) tma_load_2d(dst, map, full_bar + stage, c0, c1);
            else tma_load_2d_2sm(dst, map, full_bar + stage, c0, c1);
          };
          if (NP == 2) {
            // A is shared with the twin CTA (same rank, other pair): fetch ONE of the two boxes, multicast to both
            const uint16_t twins = (uint16_t)((1u << rank) | (1u << (rank + 2)));
            if (!(p.dbg & 8) && issuer) {
              if (!A_MN) tma_load_2d_2sm_mc(sa + pair * (BLOCK_M * 128), &tma_a, full_bar + stage, k0 + (int)pair * 64, m0, twins);
              else tma_load_2d_2sm_mc(sa + pair * (BK * 128), &tma_a, full_bar + stage, m0 + (int)pair * 64, k0, twins);
            }
          } else if (!A_MN) {
#pragma unroll
            for (int a = 0; a < KA; ++a) load(sa + a * (BLOCK_M * 128), &tma_a, k0 + a * 64, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) load(sa + i * (BK * 128), &tma_a, m0 + i * 64, k0);
          }
          if (!B_MN) {
#pragma unroll
            for (int a = 0; a < KA; ++a) load(sb + a * (BNL * 128), &tma_b, k0 + a * 64, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BNL / 64; ++i) load(sb + i * (BK * 128), &tma_b, n0 + i * 64, k0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    // The WHOLE warp runs this loop with warp-uniform values (tile, stage, descriptors) and one elected lane issues
    // the tcgen05 instructions: the operands then live in uniform registers.  (Issued from a `lane == 0` branch the
    // same code needed an ELECT / R2UR waterfall around every MMA -- ~19 dependent instructions per 64-clock MMA,
    // which made instruction issue, not the tensor core, the pace of the main loop.)
    if (leader) {
      // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, a_major bit15, b_major bit16,
      // N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BLOCK_M * CG) >> 4) << 24);
      // shared-memory descriptors (see make_smem_desc): the high word is constant, the low word is
      // (address >> 4) | (LBO >> 4) << 16 and advances by a compile-time constant per K step
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t a_lbo = A_MN ? (((uint32_t)(BK * 128) >> 4) << 16) : 0u;
      constexpr uint32_t b_lbo = B_MN ? (((uint32_t)(BK * 128) >> 4) << 16) : 0u;
      // (& 0x3FFF: inside a cluster the shared-window address carries the CTA rank above bit 24)
      const uint32_t a_lo0 = ((smem_u32(smem_a) >> 4) & 0x3FFFu) | a_lbo;
      const uint32_t b_lo0 = ((smem_u32(smem_b) >> 4) & 0x3FFFu) | b_lbo;
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = tile0; t < num_tiles; t += tile_step, ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tmem_empty + as, aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int kb_count = min(num_k_blocks, (t / mn_tiles + 1) * p.kb_per_split) - (t / mn_tiles) * p.kb_per_split;
        for (int kb = 0; kb < kb_count; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tcgen05_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * (A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * (B_BYTES >> 4);
          if (issuer && !(p.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // K-major: 64-wide atom (k / 4), then 16 elements (32 B) inside the 128 B swizzle row.
              // MN-major: 16 K-rows of 128 B (2048 B) per step; atoms along MN are BK*128 B apart (LBO).
              const uint32_t a_off = A_MN ? (uint32_t)(k * (UMMA_K * 128)) >> 4 : (uint32_t)((k >> 2) * (BLOCK_M * 128) + (k & 3) * (UMMA_K * 2)) >> 4;
              const uint32_t b_off = B_MN ? (uint32_t)(k * (UMMA_K * 128)) >> 4 : (uint32_t)((k >> 2) * (BNL * 128) + (k & 3) * (UMMA_K * 2)) >> 4;
              const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + a_off);
              const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + b_off);
              if (CG == 1) umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (issuer) {
            // frees the smem slot (in both CTAs when paired) once these MMAs retire
            if (CG == 1) umma_commit(empty_bar + stage); else umma_commit_2sm(empty_bar + stage, all_mask);
            if (kb == kb_count - 1) {
              if (CG == 1) umma_commit(tmem_full + as); else umma_commit_2sm(tmem_full + as, pair_mask);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int etid = (warp - 2) * 32 + lane;  // 0..127
    const int rloc = q * 32 + lane;           // row of the 128-row tile held by this thread
    int iter = 0;
    int sbuf = 0;
    for (int t = tile0; t < num_tiles; t += tile_step, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      const int tt = t % mn_tiles, sp = t / mn_tiles;
      const int m0 = (tt % m_tiles) * (BLOCK_M * CG) + (int)rank * BLOCK_M;
      const int n0 = ((tt / m_tiles) * NP + (int)pair) * BN;
      mbar_wait(tmem_full + as, aphase);
      tcgen05_fence_after();
      const int row = m0 + rloc;
      const bool row_ok = row < p.M;
      if (p.tma_store) {
        // ---- coalesced path: 64-column bf16 chunks -> 128B-swizzled smem tile -> TMA store (clips M / N tails)
#pragma unroll 1
        for (int c64 = 0; c64 < BN; c64 += 64) {
          if (n0 + c64 >= p.N) break;  // uniform
          uint8_t* stg = smem_st + sbuf * STG_BYTES;
          if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // staging tile sbuf is free again
          asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            const int c0 = c64 + hf * 32;
            __syncwarp();
            if (!(p.dbg & 2)) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), r);
            else { for (int j = 0; j < 32; ++j) r[j] = 0; }
            const int n = n0 + c0;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias != nullptr && n < p.N) {
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const f8 bb = load8(p.bias + n + j);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[j + e] += bb.v[e];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) v[j] += __bfloat162float(p.bias[n + j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16(v[8 * j], v[8 * j + 1]);
              u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
              u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
              u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
              const int ci = hf * 4 + j;  // 16-byte chunk inside the 128-byte row
              *reinterpret_cast<uint4*>(stg + rloc * 128 + ((ci ^ (rloc & 7)) << 4)) = u;
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (etid == 0 && !(p.dbg & 1)) {
            tma_store_2d(&tma_d, stg, n0 + c64, m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          sbuf ^= 1;
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          __syncwarp();
          if (!(p.dbg & 2)) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), r);
          else { for (int j = 0; j < 32; ++j) r[j] = 0; }
          const int n = n0 + c0;
          if (n < p.N && row_ok && !(p.dbg & 1)) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias != nullptr) {
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const f8 bb = load8(p.bias + n + j);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[j + e] += bb.v[e];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) v[j] += __bfloat162float(p.bias[n + j]);
              }
            }
            if (p.residual != nullptr) {
              const float* rp = p.residual + (int64_t)row * p.ldr + n;
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 x = *reinterpret_cast<const float4*>(rp + j);
                  v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) v[j] += rp[j];
              }
            }
            if (p.d_bf16) {
              bf16* dp = reinterpret_cast<bf16*>(p.D) + (int64_t)row * p.ldd + n;
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 u;
                  u.x = pack_bf16(v[j], v[j + 1]);
                  u.y = pack_bf16(v[j + 2], v[j + 3]);
                  u.z = pack_bf16(v[j + 4], v[j + 5]);
                  u.w = pack_bf16(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(dp + j) = u;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) dp[j] = __float2bfloat16(v[j]);
              }
            } else {
              float* dp = reinterpret_cast<float*>(p.D) + (int64_t)sp * p.split_stride + (int64_t)row * p.ldd + n;
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) dp[j] = v[j];
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(tmem_empty + as);
        else mbar_arrive_remote(tmem_empty + as, pair_leader);  // the MMA issuer waits on its pair leader's barrier
      }
    }
    if (p.tma_store && etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores landed before smem dies
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  // The driver entry point needs a current context on the CALLING thread.  A thread that has only ever used cached
  // allocations (e.g. a fresh autograd worker whose first CUDA work is this GEMM) has none bound yet:
  // cuTensorMapEncodeTiled then fails with CUDA_ERROR_INVALID_CONTEXT.  cudaFree(0) binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of stride ld elements.
int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    ofab_set_error("ofab_gemm_bf16: cuTensorMapEncodeTiled not available from the driver");
    return OFAB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ofab_set_error("ofab_gemm_bf16: cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, ptr,
                   (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return OFAB_ERR_CUDA;
  }
  return OFAB_OK;
}

template <int BN, bool A_MN, bool B_MN, int CG, int CL, int BK = (CG == 2 ? 128 : 64)>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, cudaStream_t st) {
  constexpr int NP = CL / CG;
  constexpr int stage_bytes = BLOCK_M * BK * 2 + (BN / CG) * BK * 2;
  constexpr int kFixed = 2 * BLOCK_M * 128 /*store staging*/ + 1024 /*barriers*/ + 1024 /*alignment slack*/;
  constexpr int STAGES = ((kSmemMax - kFixed) / stage_bytes) > 8 ? 8 : ((kSmemMax - kFixed) / stage_bytes);
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  constexpr int smem_bytes = STAGES * stage_bytes + kFixed;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, CG, BK, STAGES, CL>;
  static bool configured = false;
  static int max_clusters = 0;  // co-resident clusters of this shape (4-CTA clusters must fit inside a GPC)
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ofab_pdl_enabled() ? 2 : 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_gemm_bf16: cudaFuncSetAttribute");
    max_clusters = ofab_sm_count() / CL;
    if (CL > 2) {
      cfg.gridDim = dim3(max_clusters * CL);
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_gemm_bf16: cudaOccupancyMaxActiveClusters");
      if (n < 1) {
        ofab_set_error("ofab_gemm_bf16: no %d-CTA cluster of this kernel fits on the device", CL);
        return OFAB_ERR_CUDA;
      }
      if (n < max_clusters) max_clusters = n;
    }
    configured = true;
  }
  const int m_tiles = (p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG), n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * (n_tiles / NP) * p.splits;  // work items of one cluster: NP N-adjacent tiles
  const int grid = (tiles < max_clusters ? tiles : max_clusters) * CL;
  cfg.gridDim = dim3(grid);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, td, p);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_gemm_bf16 launch");
  return OFAB_OK;
}

// splits > 1: split-K into `splits` fp32 partial slabs at D (see GemmParams); forces the 256 x 128 pair tile.
int gemm_run(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
             const void* bias, const float* residual, int64_t ldr, void* D, int64_t ldd, int d_dt, int splits, int64_t split_stride,
             ofab_stream_t stream) {
  OFAB_REQUIRE(M > 0 && N > 0 && K > 0, "ofab_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  OFAB_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "ofab_gemm_bf16: dimension overflow");
  OFAB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "ofab_gemm_bf16: lda=%lld ldb=%lld must be multiples of 8 (16-byte TMA strides)", (long long)lda, (long long)ldb);
  OFAB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)D & 15) == 0, "ofab_gemm_bf16: A/B/D must be 16-byte aligned");
  OFAB_REQUIRE(ldd % (d_dt == OFAB_BF16 ? 8 : 4) == 0, "ofab_gemm_bf16: ldd=%lld must keep rows 16-byte aligned", (long long)ldd);
  OFAB_REQUIRE(residual == nullptr || (ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0), "ofab_gemm_bf16: residual must be 16-byte aligned, ldr %% 4 == 0");
  OFAB_REQUIRE(bias == nullptr || ((uintptr_t)bias & 15) == 0, "ofab_gemm_bf16: bias must be 16-byte aligned");
  OFAB_REQUIRE(d_dt == OFAB_BF16 || d_dt == OFAB_F32, "ofab_gemm_bf16: bad d_dt");
  OFAB_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K) && ldd >= N, "ofab_gemm_bf16: leading dimension smaller than the row");

  // Tile selection.  CTA pairs (cta_group::2, 256 x BN tiles) halve shared-memory traffic per MMA and are
  // preferred; BN and the 1-CTA form are chosen by wave quantisation (time ~ waves x tile width, with the
  // 1-CTA form derated for its smem-bound main loop).
  const int sms = ofab_sm_count();
  auto cost = [&](int bn, int cg) {
    const int64_t mt = (M + BLOCK_M * cg - 1) / (BLOCK_M * cg);
    const int64_t tiles = mt * ((N + bn - 1) / bn);
    const int64_t units = sms / cg;
    const double waves = (double)((tiles + units - 1) / units);
    // per-tile efficiency factors fitted to the measured sweep (profiles/r01_gemm_sweep_v4.json)
    // (128 x 64 tiles exist for completeness and tests: measured slower than 256 x 128 pair tiles even on the
    // 768 x 768 wgrads they spread over 2x the SMs -- the K loop, not the SM count, bounds those; see split-K)
    const double eff = cg == 2 ? (bn == 256 ? 1.0 : 1.25) : (bn == 256 ? 1.1 : bn == 128 ? 1.3 : 4.0);
    return waves * bn * eff;
  };
  int BN = 256, CG = 2;
  double best = 1e30;
  const int bns[3] = {256, 128, 64};
  for (int cg = 2; cg >= 1; --cg)
    for (int bi = 0; bi < 3; ++bi) {
      const int bn = bns[bi];
      if (bn == 64 && cg == 2) continue;                // 64-wide tiles only in the 1-CTA form
      if (b_mn_major && (bn / cg) % 64 != 0) continue;  // MN-major B is staged in 64-wide atoms
      const double c = cost(bn, cg);
      if (c < best - 1e-9) { best = c; BN = bn; CG = cg; }
    }
  if (const char* ov = getenv("OFAB_GEMM_BN")) {  // development overrides for tile-shape experiments
    const int v = atoi(ov);
    if (v == 64 || v == 128 || v == 256) BN = v;
  }
  if (const char* ov = getenv("OFAB_GEMM_CG")) {
    const int v = atoi(ov);
    if (v == 1 || v == 2) CG = v;
  }
  if (splits > 1) {
    BN = 128;
    CG = 2;
  }
  if (BN == 64 || (b_mn_major && (BN / CG) % 64 != 0)) CG = 1;

  CUtensorMap ta, tb;
  int rc;
  // K-major operands are loaded as 64-wide (128 B) K atoms x rows; MN-major operands as 64-wide MN atoms x BK rows
  // CTA pairs: 64-deep stages (six of them) by default -- measured 2.7 % less GEMM time over the headline step's shapes than
  // three 128-deep stages (profiles/r02_gemm_shapes_b64.txt vs r02_gemm_shapes_b64_bk64.txt: the deeper ring keeps the TMA
  // latency covered on the long-K shapes); split-K work items keep 128-deep stages (their k-block bookkeeping is in units of 128).
  uint32_t BK = (CG == 2 && splits > 1) ? 128 : 64;
  if (const char* ov = getenv("OFAB_GEMM_BK")) {  // development / tests: force the stage depth of CTA pairs
    if (atoi(ov) == 128 && CG == 2) BK = 128;
    if (atoi(ov) == 64 && CG == 2 && splits <= 1) BK = 64;
  }
  if (!a_mn_major) rc = make_map(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BLOCK_M);
  else rc = make_map(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)(BN / CG));
  else rc = make_map(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  // bf16 outputs without a residual leave through swizzled smem + TMA stores of 128 x 64 tiles (coalesced 128 B rows)
  // (N % 8: only whole 16-byte chunks are sent through the bulk-tensor store; ragged N takes the direct path)
  const int tma_store = (d_dt == OFAB_BF16 && residual == nullptr && N % 8 == 0 && splits <= 1) ? 1 : 0;
  int tma_store_f = tma_store;
  if (getenv("OFAB_GEMM_FORCE_TMA_STORE") && d_dt == OFAB_BF16 && residual == nullptr) tma_store_f = 1;  // development probe
  CUtensorMap td;
  if (tma_store_f) {
    rc = make_map(&td, D, (uint64_t)N, (uint64_t)M, (uint64_t)ldd, 64, BLOCK_M);
    if (rc) return rc;
  } else {
    td = ta;  // unused
  }

  GemmParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.bias = (const bf16*)bias;
  p.residual = residual;
  p.ldr = ldr;
  p.D = D;
  p.ldd = ldd;
  p.d_bf16 = d_dt == OFAB_BF16;
  p.tma_store = tma_store_f;
  {
    const int nkb = (int)((K + BK - 1) / BK);
    p.splits = splits > 1 ? splits : 1;
    p.kb_per_split = (nkb + p.splits - 1) / p.splits;
    p.splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;  // no empty K range
    p.split_stride = split_stride;
    OFAB_REQUIRE(splits <= 1 || p.splits == splits, "ofab_gemm_bf16: split-K plan (%d ranges) does not tile %d k-blocks", splits, nkb);
  }
  p.dbg = 0;
  if (const char* ov = getenv("OFAB_GEMM_DBG")) p.dbg = atoi(ov);
  cudaStream_t st = (cudaStream_t)stream;
  // 4-CTA clusters (two CTA pairs on N-adjacent tiles share A by TMA multicast) are correct but MEASURED SLOWER than
  // independent pairs on every shape of the benchmark step (profiles/r01_gemm_cluster4_vs_pairs.txt: +5..20 %, wgrads
  // up to +70 %): the two pairs advance in lock-step through shared empty barriers and fewer clusters are co-resident.
  // They stay available (OFAB_GEMM_CL=4, even N-tile counts) for experiments and are covered by the tests.
  int CL = CG;
  if (const char* ov = getenv("OFAB_GEMM_CL")) {
    if (atoi(ov) == 4 && CG == 2 && BK == 128 && ((N + BN - 1) / BN) % 2 == 0) CL = 4;
  }
#define GO(BNV, AM, BM)                                                      \
  do {                                                                       \
    if (CL == 4) return launch<BNV, AM, BM, 2, 4>(ta, tb, td, p, st);        \
    if (CG == 2 && BK == 64) return launch<BNV, AM, BM, 2, 2, 64>(ta, tb, td, p, st); \
    if (CG == 2) return launch<BNV, AM, BM, 2, 2>(ta, tb, td, p, st);        \
    return launch<BNV, AM, BM, 1, 1>(ta, tb, td, p, st);                     \
  } while (0)
  if (BN == 256) {
    if (!a_mn_major && !b_mn_major) GO(256, false, false);
    if (!a_mn_major && b_mn_major) GO(256, false, true);
    if (a_mn_major && !b_mn_major) GO(256, true, false);
    GO(256, true, true);
  } else if (BN == 128) {
    if (!a_mn_major && !b_mn_major) GO(128, false, false);
    if (!a_mn_major && b_mn_major) GO(128, false, true);
    if (a_mn_major && !b_mn_major) GO(128, true, false);
    GO(128, true, true);
  } else {
    if (!a_mn_major && !b_mn_major) return launch<64, false, false, 1, 1>(ta, tb, td, p, st);
    if (!a_mn_major && b_mn_major) return launch<64, false, true, 1, 1>(ta, tb, td, p, st);
    if (a_mn_major && !b_mn_major) return launch<64, true, false, 1, 1>(ta, tb, td, p, st);
    return launch<64, true, true, 1, 1>(ta, tb, td, p, st);
  }
#undef GO
}

// out[r, c] = sum_s ws[s, r, c]  (fp32 slabs [M, N]) -> D (bf16 or fp32, leading dimension ldd)
template <typename TO>
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t slab, int64_t M, int64_t N, TO* __restrict__ D,
                                     int64_t ldd) {
  pdl_launch();
  pdl_wait();
  const int64_t nvec = M * (N / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (N / 4), c = (i % (N / 4)) * 4;
    const float* src = ws + r * N + c;
    float4 acc = *reinterpret_cast<const float4*>(src);
    int sp = 1;
    for (; sp + 4 <= splits; sp += 4) {  // four slabs in flight (the adds keep the slab order: the sum does not depend on the unrolling)
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(src + (int64_t)(sp + k) * slab);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
      }
    }
    for (; sp < splits; ++sp) {
      const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)sp * slab);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (sizeof(TO) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(D) + r * ldd + c) = acc;
    } else {
      uint2 u;
      u.x = pack_bf16(acc.x, acc.y);
      u.y = pack_bf16(acc.z, acc.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(D) + r * ldd + c) = u;
    }
  }
}

// Split-K plan for a GEMM whose output has too few 256 x 128 tiles to fill the chip (wgrads of narrow layers: the
// tile count does not grow with the batch, the K loop does).  Model fitted to measurements on B200: one k-block
// (128 deep) of a pair tile takes ~0.58 us (+4 us per launch); splitting costs ~8 us (second launch, partial-slab
// round trip) plus the reduction's read of `splits` fp32 slabs at ~4 TB/s.  Returns 1 when splitting does not pay.
int splitk_plan(int64_t M, int64_t N, int64_t K) {
  if (N % 8 != 0) return 1;
  const int64_t nkb = (K + 127) / 128;
  // ranges of equal k-block count with no empty range: `want` ranges -> ceil(nkb / ceil(nkb / want)) ranges
  auto effective = [nkb](int64_t want) {
    const int64_t per = (nkb + want - 1) / want;
    return (int)((nkb + per - 1) / per);
  };
  if (const char* ov = getenv("OFAB_GEMM_SPLITS")) {  // development / test override
    const int v = atoi(ov);
    if (v >= 1 && v <= 16) return effective(v < nkb ? v : nkb);
  }
  if (nkb < 16) return 1;  // short contractions: nothing to win
  const int64_t tiles = ((M + 255) / 256) * ((N + 127) / 128);
  const int64_t units = ofab_sm_count() / 2;
  int best = 1;
  double best_t = 1e30;
  const int cand[6] = {1, 2, 3, 4, 6, 8};
  for (int ci = 0; ci < 6; ++ci) {
    const int sp = cand[ci];
    if (sp > 1 && nkb / sp < 4) break;
    const double waves = (double)((tiles * sp + units - 1) / units);
    const double t_gemm = waves * (double)((nkb + sp - 1) / sp) * 0.58 + 4.0;
    const double t_red = sp > 1 ? 8.0 + (sp + 0.5) * (double)M * (double)N * 4.0 / 4e6 : 0.0;
    if (t_gemm + t_red < best_t - 1e-9) {
      best_t = t_gemm + t_red;
      best = sp;
    }
  }
  return effective(best);
}

}  // namespace

extern "C" int ofab_gemm_bf16(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, int a_mn_major, const void* B,
                              int64_t ldb, int b_mn_major, const void* bias, const float* residual, int64_t ldr, void* D,
                              int64_t ldd, int d_dt, ofab_stream_t stream) {
  return gemm_run(M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, bias, residual, ldr, D, ldd, d_dt, 1, 0, stream);
}

extern "C" int64_t ofab_gemm_splitk_workspace_elems(int64_t M, int64_t N, int64_t K) {
  const int sp = (M > 0 && N > 0 && K > 0) ? splitk_plan(M, N, K) : 1;
  return sp > 1 ? (int64_t)sp * M * N : 0;
}

extern "C" int ofab_gemm_bf16_splitk(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, int a_mn_major, const void* B,
                                     int64_t ldb, int b_mn_major, void* D, int64_t ldd, int d_dt, float* workspace,
                                     int64_t workspace_elems, ofab_stream_t stream) {
  OFAB_REQUIRE(M > 0 && N > 0 && K > 0, "ofab_gemm_bf16_splitk: empty problem");
  const int sp = splitk_plan(M, N, K);
  if (sp <= 1) return gemm_run(M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, nullptr, nullptr, 0, D, ldd, d_dt, 1, 0, stream);
  OFAB_REQUIRE(workspace != nullptr && workspace_elems >= (int64_t)sp * M * N && ((uintptr_t)workspace & 15) == 0,
               "ofab_gemm_bf16_splitk: workspace too small (need ofab_gemm_splitk_workspace_elems = %lld fp32) or unaligned",
               (long long)((int64_t)sp * M * N));
  OFAB_REQUIRE(d_dt == OFAB_BF16 || d_dt == OFAB_F32, "ofab_gemm_bf16_splitk: bad d_dt");
  OFAB_REQUIRE(ldd >= N && ldd % 4 == 0 && ((uintptr_t)D & 15) == 0, "ofab_gemm_bf16_splitk: D must be 16-byte aligned with ldd %% 4 == 0");
  int rc = gemm_run(M, N, K, A, lda, a_mn_major, B, ldb, b_mn_major, nullptr, nullptr, 0, workspace, N, OFAB_F32, sp, M * N, stream);
  if (rc) return rc;
  const int64_t nvec = M * (N / 4);
  const int grid = (int)((nvec + 255) / 256 < (int64_t)ofab_sm_count() * 8 ? (nvec + 255) / 256 : (int64_t)ofab_sm_count() * 8);
  if (d_dt == OFAB_F32)
    ofab_launch((splitk_reduce_kernel<float>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, workspace, sp, M * N, M, N, (float*)D, ldd);
  else
    ofab_launch((splitk_reduce_kernel<bf16>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, workspace, sp, M * N, M, N, (bf16*)D, ldd);
  OFAB_LAUNCH_CHECK("ofab_gemm_bf16_splitk reduce");
  return OFAB_OK;
}
