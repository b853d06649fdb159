// DRAFT -- NOT BUILT INTO libofab, NOT WIRED, NEVER RUN ON A GPU.  Round-2 starting point for DESIGN.md Appendix A
// ("single-pass tcgen05 attention").  It only has to compile (nvcc / ptxas accept every instruction form used here:
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -c tools/drafts/attn_tc_fwd_draft.cu -o /dev/null );
// every layout assumption that has not been exercised on hardware is marked UNVERIFIED.
//
// Forward of the plain variant (no position columns / table), head_dim 64, all keys resident (Tk_pad <= 320 here):
//   one CTA = one (b, h, 128 query rows);  S = Q K^T -> TMEM;  softmax straight from TMEM (one thread per query row,
//   no online rescaling);  P (bf16) written back over S;  O = P V with A read from TMEM and V MN-major from shared memory.
#include <cuda.h>

#include "../../ofasys_b200/csrc/common.cuh"

namespace {

constexpr int BM = 128;         // query rows per work item (UMMA M)
constexpr int TK_MAX = 320;     // keys resident in this draft: encoder self 265, decoder cross 265
constexpr int kThreads = 192;   // warp 0: TMA + MMA issuer, warp 1: TMEM allocator, warps 2..5: softmax / epilogue
constexpr float kLog2e = 1.4426950408889634f;

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
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (A operand read from tensor memory: the probabilities)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct TcAttnParams {
  int B, H, Tq, Tk;
  int q_col0, k_col0, v_col0;  // first column of head 0's q / k / v inside the packed projection rows
  float scale;
  const uint8_t* kpm;          // [B, Tk] or NULL
  int causal;
  bf16* o;
  int64_t o_bs, o_rs;
  float* lse;                  // [B, H, Tq]
};

// Instruction descriptor of kind::f16 (as in gemm.cu): D = f32 (bit 4), A = B = bf16 (bits 7, 10), b_major at bit 16,
// N >> 3 at [17, 23), M >> 4 at [24, 29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory descriptor, SWIZZLE_128B, 8-row groups 1024 B apart (gemm.cu: desc_hi)
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) { return ((uint64_t)kDescHi << 32) | (uint64_t)((saddr >> 4) & 0x3FFFu); }

// map_q: rows of the packed q source ([B * Tq, ld]), box 64 columns x 128 rows; map_kv: rows of the k|v source, box 64 x 64.
__global__ void __launch_bounds__(kThreads, 1) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                                                                  const TcAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                        // 128 rows x 128 B
  uint8_t* sK = sQ + BM * 128;               // TK_MAX rows x 128 B (K-major: a row is one key's 64 values)
  uint8_t* sV = sK + TK_MAX * 128;           // TK_MAX rows x 128 B (the same bytes read MN-major: N = head_dim contiguous)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + TK_MAX * 128);  // [0] loads, [1] S ready, [2] P written, [3] O ready
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  uint32_t* kmask = tmem_ptr + 2;            // TK_MAX / 32 words: bit j set <=> key j is in range and not padding

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BM;
  const int tk_pad = (p.Tk + 15) & ~15;      // UMMA N granularity for M = 128
  const int n_kbox = (tk_pad + 63) >> 6;     // 64-row TMA boxes of K / V

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_init(bars + 2, 128);  // every softmax thread arrives once its row of P is in TMEM
    mbar_init(bars + 3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // key-validity bitmap (as build_kmask_all in attn.cu)
  for (int w0 = warp; w0 < TK_MAX / 32; w0 += kThreads / 32) {
    const int j = w0 * 32 + lane;
    bool ok = j < p.Tk;
    if (ok && p.kpm != nullptr) ok = p.kpm[(int64_t)b * p.Tk + j] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) kmask[w0] = m;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;         // fp32 scores, columns [0, tk_pad); later bf16 P in columns [0, tk_pad / 2)
  const uint32_t tmem_O = tmem_base + 448;   // fp32 output accumulator, 64 columns

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer + MMA issuer (one elected lane)
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(bars + 0, (uint32_t)(BM * 128 + 2 * n_kbox * 64 * 128));
      tma_load_2d(sQ, &map_q, bars + 0, p.q_col0 + h * 64, b * p.Tq + q0);
      for (int i = 0; i < n_kbox; ++i) {
        tma_load_2d(sK + i * 64 * 128, &map_kv, bars + 0, p.k_col0 + h * 64, b * p.Tk + i * 64);
        tma_load_2d(sV + i * 64 * 128, &map_kv, bars + 0, p.v_col0 + h * 64, b * p.Tk + i * 64);
      }
    }
    __syncwarp();
    mbar_wait(bars + 0, 0);
    tc_fence_after();
    if (issuer) {
      // S = Q K^T: M = 128, N in chunks of <= 256 (multiples of 16), K = 64 = 4 steps of 16 (32 B inside the 128 B row)
      for (int n0 = 0; n0 < tk_pad; n0 += 256) {
        const int n = min(256, tk_pad - n0);
        const uint32_t idesc = make_idesc(BM, n, false);
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_S + n0, smem_desc(smem_u32(sQ) + k * 32), smem_desc(smem_u32(sK) + n0 * 128 + k * 32), idesc, k != 0);
      }
      umma_commit(bars + 1);
    }
    __syncwarp();
    mbar_wait(bars + 2, 0);  // P is in TMEM
    tc_fence_after();
    if (issuer) {
      // O = P V: A = P from TMEM (bf16 pairs: 16 keys = 8 columns per step  -- UNVERIFIED packing), B = V MN-major
      // (N = 64 contiguous, 16 key rows = 2048 B per step, as the B_MN operand of gemm.cu), M = 128, N = 64
      const uint32_t idesc = make_idesc(BM, 64, true);
      for (int ks = 0; ks < tk_pad / 16; ++ks)
        umma_ts(tmem_O, tmem_S + ks * 8, smem_desc(smem_u32(sV) + ks * 2048), idesc, ks != 0);
      umma_commit(bars + 3);
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ softmax + epilogue: one thread per query row
    const int quad = warp & 3;                      // a warp may touch TMEM lanes [32 * (warp % 4), +32)
    const int row = quad * 32 + lane;               // query row inside the tile == TMEM lane
    const int i = q0 + row;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float c2 = p.scale * kLog2e;
    mbar_wait(bars + 1, 0);
    tc_fence_after();
    // pass 1: row maximum over the valid keys
    float m2 = -INFINITY;
    for (int c0 = 0; c0 < tk_pad; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_S + lane_addr + c0, r);
      const uint32_t km = kmask[c0 >> 5];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool ok = ((km >> j) & 1u) && (!p.causal || c0 + j <= i);
        if (ok) m2 = fmaxf(m2, __uint_as_float(r[j]) * c2);
      }
    }
    const float m_use = m2 == -INFINITY ? 0.f : m2;
    // pass 2: probabilities (log2 domain), row sum, bf16 pairs written back over the scores already consumed
    float l = 0.f;
    for (int c0 = 0; c0 < tk_pad; c0 += 32) {
      uint32_t r[32], pk[16];
      tmem_ld32(tmem_S + lane_addr + c0, r);
      const uint32_t km = kmask[c0 >> 5];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const bool ok0 = ((km >> j) & 1u) && (!p.causal || c0 + j <= i);
        const bool ok1 = ((km >> (j + 1)) & 1u) && (!p.causal || c0 + j + 1 <= i);
        const float p0 = ok0 ? fast_ex2(fmaf(__uint_as_float(r[j]), c2, -m_use)) : 0.f;
        const float p1 = ok1 ? fast_ex2(fmaf(__uint_as_float(r[j + 1]), c2, -m_use)) : 0.f;
        l += p0 + p1;
        pk[j >> 1] = pack_bf16(p0, p1);
      }
      tmem_st16(tmem_S + lane_addr + (c0 >> 1), pk);  // columns [c0/2, c0/2 + 16) <= columns already read
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bars + 2);
    // epilogue: O / l
    mbar_wait(bars + 3, 0);
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    if (i < p.Tq) {
      bf16* op = p.o + (int64_t)b * p.o_bs + (int64_t)i * p.o_rs + h * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(tmem_O + lane_addr + half * 32, r);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          f8 v;
#pragma unroll
          for (int e = 0; e < 8; ++e) v.v[e] = __uint_as_float(r[j + e]) * inv;
          store8(op + half * 32 + j, v);
        }
      }
      p.lse[((int64_t)b * p.H + h) * p.Tq + i] = l > 0.f ? (m2 + __log2f(l)) * 0.6931471805599453f : -INFINITY;
    } else {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_addr, r);  // keep the warp's tcgen05.ld collective (all 32 lanes participate)
      tmem_ld32(tmem_O + lane_addr + 32, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

}  // namespace

// host side of the draft: shared-memory size and grid (tensor-map creation as make_map in gemm.cu: bf16, 2-D, box {64, 128}
// for Q and {64, 64} for K|V, SWIZZLE_128B)
extern "C" int ofab_draft_attn_tc_fwd_smem(void) { return BM * 128 + 2 * TK_MAX * 128 + 4 * 8 + 8 + (TK_MAX / 32) * 4 + 1024; }
