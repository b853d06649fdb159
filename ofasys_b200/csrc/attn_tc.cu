// Fused attention on the 5th-generation tensor cores (tcgen05) for head_dim 64: forward, dQ and dK/dV kernels.
//
// Replaces MultiheadAttention.forward's bmm / +bias / mask / softmax / dropout / bmm chain
// (ofasys/module/multihead_attention.py:308-338) and its autograd backward.  Scores and probabilities never leave the
// SM: S = Q K^T is accumulated in tensor memory (TMEM), the softmax warps read it with tcgen05.ld (one thread per
// row), write the bf16 probabilities back OVER the scores with tcgen05.st, and the second product (P V, dS K, P^T dO,
// dS^T Q) takes its A operand straight from TMEM -- no shared-memory round trip and no transposed copies: the same
// 128-byte-swizzled shared-memory tile is read K-major by the score-type products and MN-major by the second ones
// (the descriptors of gemm.cu).
//
// The position bias of OFA's "Mode A" (adaptor/general.py:223-282: abs-pos term + per-slot relative-position tables) is a
// dense additive tile bias[h, i, j] in fp16 that does not depend on the batch index (the reference materialises
// [B, H, S, S] per layer); the kernels add it to the scores; the dQ kernel emits dS so that ONE reduction per layer
// turns it into the table / position-projection gradients (ofab_attn_bias_bwd).
//
// Work split (all three kernels): CTA = one (batch, head, 128-row tile); 192 threads:
//   warp 0      TMA producer (cp.async.bulk.tensor.3d, [cols, T, B] maps: rows past T are zero-filled per batch)
//   warp 1      MMA issuer (one elected lane, warp-uniform control flow) + TMEM allocation (256 columns -> 2 CTAs / SM)
//   warps 2..5  softmax / gradient math, one TMEM lane (= row) per thread, 32-column chunks
// Rows = queries in the forward and dQ kernels, keys in the dK/dV kernel (which works on S^T so that dK / dV
// accumulate in TMEM over its loop over query blocks: no atomics).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------- PTX wrappers (as gemm.cu)
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
// Bounded wait: a protocol bug must not hang the GPU -- after ~2^22 failed polls (seconds; legitimate waits are microseconds)
// the kernel traps and the launch reports an error.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  int spins = 0;
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
    if (!ok && ++spins > (1 << 22)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
// D[tmem] (+)= A[smem] * B[smem]
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
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (bf16 pairs, one 32-bit column per two K elements) is read from tensor memory
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld32_nowait(taddr, r);
  tmem_ld_wait();
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Instruction descriptor of kind::f16 (gemm.cu): D = f32 (bit 4), A = B = bf16 (bits 7, 10), b_major at bit 16,
// N >> 3 at [17, 23), M >> 4 at [24, 29).  M = 128 always here (one TMEM lane per row).
__device__ __forceinline__ uint32_t make_idesc(int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// Shared-memory descriptor, SWIZZLE_128B, 8-row groups 1024 B apart (gemm.cu: desc_hi).  Tiles are [rows, 128 B]:
//   K-major  (rows = M or N index, the 64 head-dim values contiguous): k-step of 16 elements = +32 B inside the row
//   MN-major (rows = K index, the 64 head-dim values are the N index): k-step of 16 rows = +2048 B; one 64-wide atom
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo_bytes = 0) {
  return ((uint64_t)kDescHi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16));
}

// ---- attention dropout: the keyed hash of attn.cu (same bits for the same (seed, step, site, b, h, i, j)) --------------
struct TcDrop {
  const unsigned long long* state;
  uint32_t site, thresh32;
  float inv_keep;
};
__device__ __forceinline__ uint2 tc_drop_key(const TcDrop& d, int bh) {
  const unsigned long long seed = d.state[0], step = d.state[1];
  const uint4 r = philox4x32_7(make_uint4((uint32_t)bh, d.site, (uint32_t)step, 0x6A09E667u),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
  return make_uint2(r.x, r.y | 1u);
}
__device__ __forceinline__ bool tc_keep(const uint2 key, uint32_t ij, uint32_t thresh32) {
  uint32_t x = (ij ^ key.x) * 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= key.y;
  x ^= x >> 16;
  x *= 0xC2B2AE3Du;
  x ^= x >> 15;
  return x >= thresh32;
}

struct TcParams {
  int B, H, Tq, Tk;
  float scale;
  const uint8_t* kpm;     // [B, Tk] or NULL
  int causal;
  const __half* bias;     // [H, Tq, bias_ld] (rows = queries) or NULL
  const __half* bias_t;   // [H, Tk, bias_t_ld] (rows = keys: the transposed tile, dK/dV kernel) or NULL
  int64_t bias_hs, bias_t_hs;
  int bias_ld, bias_t_ld;
  // forward outputs / saved
  bf16* o;
  int64_t o_bs, o_rs;
  float* lse;             // [B, H, Tq]
  // backward
  const bf16* d_o;
  int64_t do_bs, do_rs;
  float* delta;           // [B, H, Tq]
  bf16 *dq, *dk, *dv;
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  float *dq_colsum, *dk_colsum, *dv_colsum;  // [B * tiles, H * 64] per-CTA column sums or NULL
  bf16* ds;               // [B, H, Tq, bias_ld] raw dS = P o (dP - delta) (the gradient of the additive bias, per batch) or NULL
  TcDrop drop;
  int hpc;                // forward: heads per CTA (the CTA works on heads [blockIdx.y * hpc, +hpc) one after the other)
};

// Key-validity bitmap: bit j of word j / 32 set <=> key j < Tk and not padding; `words` 32-bit words are written.
__device__ __forceinline__ void build_kmask(const TcParams& p, int b, uint32_t* dst, int words) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w0 = warp; w0 < words; w0 += kThreads / 32) {
    const int j = w0 * 32 + lane;
    bool ok = j < p.Tk;
    if (ok && p.kpm != nullptr) ok = p.kpm[(int64_t)b * p.Tk + j] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) dst[w0] = m;
  }
}

// N consecutive fp16 bias values of one row starting at column c (N = 16 or 32) as raw 16-byte vectors (prefetched one chunk
// ahead and converted when used); columns >= ld read as 0
template <int NV>
__device__ __forceinline__ void load_bias_raw(const __half* __restrict__ row, int c, int ld, bool row_ok, uint4 (&u)[NV]) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    u[v] = make_uint4(0u, 0u, 0u, 0u);
    if (row_ok && c + 8 * v < ld) u[v] = __ldg(reinterpret_cast<const uint4*>(row + c + 8 * v));
  }
}
__device__ __forceinline__ float bias_at(const uint4* u, int j) {  // element j of the raw chunk (the tile is stored in the log2 domain)
  const __half* h = reinterpret_cast<const __half*>(u);
  return __half2float(h[j]);
}

// Column sums over the warp's 32 rows of v[32] (one row per lane): a transpose-reduce of 31 shuffles; afterwards lane l
// holds the sum of column l.
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = upper ? v[j] : v[j + n / 2];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      v[j] = (upper ? v[j + n / 2] : v[j]) + recv;
    }
  }
  return v[0];
}
// Per-CTA column sums of a 128 x 64 gradient tile held in TMEM (the bias gradient of the projection that produced q / k / v
// is the column sum of dq / dk / dv, nn.Linear backward of multihead_attention.py:199-218): each of the 4 compute warps
// reduces its 32 rows, the warps meet in shared memory, 64 threads write one partial row.  Called by all 128 compute threads.
__device__ __forceinline__ void tile_colsum_out(float (&lo)[32], float (&hi)[32], float* cs /* smem [4][64] */, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, quad = (threadIdx.x >> 5) & 3;
  const float a = warp_colsum32(lo), b = warp_colsum32(hi);
  cs[quad * 64 + lane] = a;
  cs[quad * 64 + 32 + lane] = b;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int t = (threadIdx.x - 64);  // compute threads are 64..191
  if (t < 64) out[t] = cs[t] + cs[64 + t] + cs[128 + t] + cs[192 + t];
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// ===================================================================================== forward
// TMEM columns: S0 | S1 (fp32 scores of the even / odd 64-key block) [0, 128) | P0 | P1 (their bf16 probabilities, the A operand
// of P V) [128, 192) | O (fp32, accumulates P V over all key blocks) [192, 256).
// The MMA warp runs one block ahead: S(kb+1) = Q K^T is issued before it waits for the probabilities of block kb, so the
// tensor core works on the next scores while the softmax warps exponentiate the current ones; P V of block kb overlaps the
// softmax of block kb+1.
// Online softmax with LAZY, OPTIMISTIC rescaling: probabilities are taken relative to a reference maximum m_ref fixed by
// the first block; later blocks are exponentiated in ONE pass against it while their maximum is tracked on the side, and
// only if that exceeds m_ref by more than 2^8 is the reference raised, the O accumulator in TMEM rescaled and the block
// redone (the scores are still in TMEM).  Otherwise the row sum and O simply carry a common factor <= 256, exact after
// the final division.  With OFA's score ranges the redo path is taken (almost) never.
template <bool MASKED, bool HAS_BIAS, bool DROP>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&r)[32], const uint4* bu, float c2, float m_use, uint32_t km, bool causal, int col0, int i,
                                          float& mx, float& ls, uint32_t (&pk)[16], const uint2 dkey, const TcDrop& drop, int Tk) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float pe[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float x = fmaf(__uint_as_float(r[j + e]), c2, -m_use);  // relative to the reference maximum
      if (HAS_BIAS) x += bias_at(bu, j + e);
      if (MASKED) {
        bool ok = (km >> (j + e)) & 1u;
        if (causal) ok = ok && (col0 + j + e <= i);
        x = ok ? x : -INFINITY;
      }
      mx = fmaxf(mx, x);
      pe[e] = fast_ex2(x);
    }
    ls += pe[0] + pe[1];
    if (DROP) {  // the row sum is that of the undropped probabilities; only P V sees the mask
      const uint32_t ij = (uint32_t)i * (uint32_t)Tk + (uint32_t)(col0 + j);
      pe[0] = tc_keep(dkey, ij, drop.thresh32) ? pe[0] * drop.inv_keep : 0.f;
      pe[1] = tc_keep(dkey, ij + 1u, drop.thresh32) ? pe[1] * drop.inv_keep : 0.f;
    }
    pk[j >> 1] = pack_bf16(pe[0], pe[1]);
  }
}

// Rare path of the forward kernel: the block's maximum exceeded the reference by more than 2^8 for some row of this warp.
// Rescale the row sum and this warp's rows of the O accumulator to the new reference and exponentiate the block again
// (its scores are still in TMEM).  Not inlined: keeps the common path's register budget small.
template <bool HAS_BIAS, bool DROP>
__device__ __noinline__ void fwd_redo_block(uint32_t tS, uint32_t tP, uint32_t tO, int nch, const __half* brow, const TcParams& p, float c2,
                                            float m_ref, float m_new, uint32_t km0, uint32_t km1, int k0, int i, bool row_ok, float& l_run,
                                            float& ls, const uint2 dkey) {
  const float corr = (m_new != m_ref) ? fast_ex2(m_ref - m_new) : 1.0f;  // m_ref = -inf -> 0 (its O row and l are 0 anyway)
#pragma unroll 1
  for (int hf = 0; hf < 2; ++hf) {
    uint32_t o[32];
    tmem_ld32(tO + hf * 32, o);
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
    tmem_st32(tO + hf * 32, o);
  }
  l_run *= corr;
  const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
  float mx = -INFINITY;
  ls = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < nch; ++ch) {
    uint32_t r[32], pk[16];
    uint4 bu[HAS_BIAS ? 4 : 1];
    if constexpr (HAS_BIAS) load_bias_raw<4>(brow, k0 + ch * 32, p.bias_ld, row_ok, bu);
    tmem_ld32(tS + ch * 32, r);
    fwd_chunk<true, HAS_BIAS, DROP>(r, bu, c2, m_use, ch == 0 ? km0 : km1, p.causal, k0 + ch * 32, i, mx, ls, pk, dkey, p.drop, p.Tk);
    tmem_st16(tP + ch * 16, pk);
  }
}

// One CTA works on `hpc` heads of one (batch, 128-query tile) one after the other: TMEM allocation, barrier set-up and the key
// bitmap are paid once, and the producer / MMA warps run ahead ACROSS heads (Q double-buffered; the next head's first scores are
// issued while the current head's last block is exponentiated), so the per-work-item prologue that dominated the one-head-per-CTA
// form (ncu: ~10 us CTA lifetime for ~3 us of math) is hidden.  `g` counts key blocks over all heads of the CTA: ring stages and
// score / probability buffers are indexed by it.
template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(kThreads, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const TcParams p) {
  constexpr int BM = 128, BN = 64, STAGES = 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // 2 x (128 x 128 B)
  uint8_t* sK = sQ + 2 * BM * 128;                 // STAGES x (64 x 128 B)
  uint8_t* sV = sK + STAGES * BN * 128;            // STAGES x (64 x 128 B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * BN * 128);
  uint64_t* q_full = bars;                         // [2]
  uint64_t* q_empty = q_full + 2;                  // [2] every S product of the head that used this Q buffer has retired
  uint64_t* full = q_empty + 2;                    // STAGES
  uint64_t* empty = full + STAGES;                 // STAGES
  uint64_t* s_full = empty + STAGES;               // [2] S(g) is in TMEM buffer g & 1
  uint64_t* p_ready = s_full + 2;                  // [2] P(g) is in TMEM buffer g & 1 (4 warp arrivals)
  uint64_t* pv_done = p_ready + 2;                 // one completion per P V product (waited for only before a rescale)
  uint64_t* o_done = pv_done + 1;                  // the head's last P V has retired
  uint64_t* o_free = o_done + 1;                   // the head's output has been read out of TMEM (4 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 1);
  uint32_t* kmask_s = tmem_ptr + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.z;
  const int h0 = blockIdx.y * p.hpc;
  const int nh = min(p.hpc, p.H - h0);             // heads of this CTA
  const int q0 = qt * BM;
  int nkb = (p.Tk + BN - 1) / BN;
  if (p.causal) nkb = min(nkb, (q0 + BM - 1) / BN + 1);  // key blocks past the last query of the tile see nothing
  const int kwords = nkb * (BN / 32);
  const int total = nh * nkb;

  pdl_launch();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full + i, 1);
      mbar_init(q_empty + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(p_ready + i, 4);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_done, 1);
    mbar_init(o_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();  // global memory (masks, TMA loads, bias, outputs) only below
  build_kmask(p, b, kmask_s, kwords);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_P = tmem_base + 128, tmem_O = tmem_base + 192;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool issuer = elect_one();
    auto load_q = [&](int hi) {  // Q tile of the CTA's head hi into buffer hi & 1 (free once head hi - 2 has issued all its S)
      mbar_wait(q_empty + (hi & 1), ((hi >> 1) & 1) ^ 1);
      if (issuer) {
        mbar_expect_tx(q_full + (hi & 1), BM * 128);
        tma_load_3d(sQ + (hi & 1) * BM * 128, &map_q, q_full + (hi & 1), (h0 + hi) * 64, q0, b);
      }
      __syncwarp();
    };
    load_q(0);
    for (int g = 0; g < total; ++g) {
      const int hi = g / nkb, kb = g - hi * nkb;
      const int st = g % STAGES;
      mbar_wait(empty + st, ((g / STAGES) & 1) ^ 1);
      if (issuer) {
        mbar_expect_tx(full + st, 2 * BN * 128);
        tma_load_3d(sK + st * BN * 128, &map_k, full + st, (h0 + hi) * 64, kb * BN, b);
        tma_load_3d(sV + st * BN * 128, &map_v, full + st, (h0 + hi) * 64, kb * BN, b);
      }
      __syncwarp();
      if (kb == 0 && hi + 1 < nh) load_q(hi + 1);  // the next head's queries arrive under this head's math
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool issuer = elect_one();
    auto issue_s = [&](int g) {  // S(g) = Q K^T into score buffer g & 1
      const int hi = g / nkb, kb = g - hi * nkb;
      const int st = g % STAGES;
      if (kb == 0) {
        mbar_wait(q_full + (hi & 1), (hi >> 1) & 1);
        tc_fence_after();
      }
      mbar_wait(full + st, (g / STAGES) & 1);
      tc_fence_after();
      if (issuer) {
        const int n = min(BN, ((p.Tk - kb * BN) + 15) & ~15);  // keys of this block rounded up to the UMMA N granularity
        const uint32_t idesc = make_idesc(n, false);
        const uint32_t q_addr = smem_u32(sQ + (hi & 1) * BM * 128), k_addr = smem_u32(sK + st * BN * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_S + (g & 1) * BN, sdesc(q_addr + k * 32), sdesc(k_addr + k * 32), idesc, k != 0);
        umma_commit(s_full + (g & 1));  // (tracks every MMA issued so far)
        if (kb == nkb - 1) umma_commit(q_empty + (hi & 1));  // the head's last use of its Q buffer
      }
      __syncwarp();
    };
    issue_s(0);
    for (int g = 0; g < total; ++g) {
      const int hi = g / nkb, kb = g - hi * nkb;
      const int st = g % STAGES;
      if (g + 1 < total) issue_s(g + 1);  // one block ahead (its buffer was freed by P(g-1), which this warp has already waited for)
      mbar_wait(p_ready + (g & 1), (g >> 1) & 1);
      if (kb == 0 && hi > 0) mbar_wait(o_free, (hi - 1) & 1);  // the previous head's output has left TMEM
      tc_fence_after();
      if (issuer) {
        const int n = min(BN, ((p.Tk - kb * BN) + 15) & ~15);
        const uint32_t idesc = make_idesc(64, true);
        const uint32_t v_addr = smem_u32(sV + st * BN * 128);
        for (int ks = 0; ks < n / 16; ++ks)
          umma_ts(tmem_O, tmem_P + (g & 1) * (BN / 2) + ks * 8, sdesc(v_addr + ks * 2048, BN * 128), idesc, (kb | ks) != 0);
        umma_commit(pv_done);
        umma_commit(empty + st);
        if (kb == nkb - 1) umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax: one thread per query row
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int i = q0 + row;
    const bool row_ok = i < p.Tq;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float c2 = p.scale * kLog2e;
    for (int hi = 0; hi < nh; ++hi) {
      const int h = h0 + hi;
      const __half* brow = HAS_BIAS ? p.bias + (int64_t)h * p.bias_hs + (int64_t)(row_ok ? i : 0) * p.bias_ld : nullptr;
      uint2 dkey = make_uint2(0u, 1u);
      if (DROP) dkey = tc_drop_key(p.drop, b * p.H + h);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int g = hi * nkb + kb;
        const int k0 = kb * BN;
        const int n = min(BN, ((p.Tk - k0) + 15) & ~15);
        const int nch = (n + 31) >> 5;  // 1 or 2 chunks of 32 columns
        const uint32_t tS = tmem_S + lane_addr + (g & 1) * BN, tP = tmem_P + lane_addr + (g & 1) * (BN / 2);
        const bool diag = p.causal && (k0 + BN - 1 > q0);  // some (i, j) of this tile may have j > i
        uint4 bu[2][HAS_BIAS ? 4 : 1];
        if constexpr (HAS_BIAS) {  // in flight while the MMAs run
          load_bias_raw<4>(brow, k0, p.bias_ld, row_ok, bu[0]);
          if (nch > 1) load_bias_raw<4>(brow, k0 + 32, p.bias_ld, row_ok, bu[1]);
        }
        const uint32_t km0 = kmask_s[k0 >> 5], km1 = kmask_s[(k0 >> 5) + 1];
        const bool masked0 = diag || km0 != 0xffffffffu, masked1 = diag || km1 != 0xffffffffu;
        mbar_wait(s_full + (g & 1), (g >> 1) & 1);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(tS, r0);
        if (nch > 1) tmem_ld32_nowait(tS + 32, r1);
        tmem_ld_wait();
        if (kb == 0) {  // the first block fixes the reference maximum
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r0[j]) * c2;
            if (HAS_BIAS) x += bias_at(bu[0], j);
            bool ok = (km0 >> j) & 1u;
            if (p.causal) ok = ok && (k0 + j <= i);
            mx = fmaxf(mx, ok ? x : -INFINITY);
          }
          if (nch > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(r1[j]) * c2;
              if (HAS_BIAS) x += bias_at(bu[1], j);
              bool ok = (km1 >> j) & 1u;
              if (p.causal) ok = ok && (k0 + 32 + j <= i);
              mx = fmaxf(mx, ok ? x : -INFINITY);
            }
          }
          // an INTEGER reference (log2 domain): every later rescaling is by an exact power of two, so the bf16 rounding of the
          // probabilities does not depend on the key-block order (the oracle's storage model reproduces it without knowing it)
          m_ref = ceilf(mx);
        }
        {
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          float mx = -INFINITY, ls = 0.f;  // mx: block maximum RELATIVE to m_use
          uint32_t pk[16];
          if (masked0) fwd_chunk<true, HAS_BIAS, DROP>(r0, bu[0], c2, m_use, km0, p.causal, k0, i, mx, ls, pk, dkey, p.drop, p.Tk);
          else fwd_chunk<false, HAS_BIAS, DROP>(r0, bu[0], c2, m_use, km0, p.causal, k0, i, mx, ls, pk, dkey, p.drop, p.Tk);
          tmem_st16(tP, pk);
          if (nch > 1) {
            if (masked1) fwd_chunk<true, HAS_BIAS, DROP>(r1, bu[1], c2, m_use, km1, p.causal, k0 + 32, i, mx, ls, pk, dkey, p.drop, p.Tk);
            else fwd_chunk<false, HAS_BIAS, DROP>(r1, bu[1], c2, m_use, km1, p.causal, k0 + 32, i, mx, ls, pk, dkey, p.drop, p.Tk);
            tmem_st16(tP + 16, pk);
          }
          // optimistic pass done: was the reference maximum still good for every row of this warp?
          const bool raise = kb > 0 && (mx > 8.0f || (m_ref == -INFINITY && mx != -INFINITY));
          if (__any_sync(0xffffffffu, raise)) {  // rare (warp-uniform: the TMEM accesses of the redo are collective)
            // every earlier P V must have retired: pv_done has completed g - 1 or g times here (s_full(g) implies P V(g-2))
            mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_after();
            const float m_new = raise ? ceilf(m_use + mx) : m_ref;
            fwd_redo_block<HAS_BIAS, DROP>(tS, tP, tmem_O + lane_addr, nch, brow, p, c2, m_ref, m_new, km0, km1, k0, i, row_ok, l_run, ls, dkey);
            m_ref = m_new;
          }
          l_run += ls;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready + (g & 1));
      }
      mbar_wait(o_done, hi & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld32(tmem_O + lane_addr + hf * 32, r);
        if (row_ok) {
          bf16* op = p.o + (int64_t)b * p.o_bs + (int64_t)i * p.o_rs + h * 64 + hf * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            f8 v;
#pragma unroll
            for (int e = 0; e < 8; ++e) v.v[e] = __uint_as_float(r[j + e]) * inv;
            store8(op + j, v);
          }
        }
      }
      tc_fence_before();  // the accumulator columns may be overwritten by the next head's first P V
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (row_ok) p.lse[((int64_t)b * p.H + h) * p.Tq + i] = l_run > 0.f ? (m_ref + __log2f(l_run)) * kLn2 : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// One 16-column chunk of the dQ kernel: dS = P o (dP - delta) with P = ex2(x - lse), x in the log2 domain.
// out_s: scale * dS (what dQ += dS K consumes); raw dS = out_s / scale is only needed for the bias gradient.
template <bool MASKED, bool HAS_BIAS, bool DROP>
__device__ __forceinline__ void dq_chunk(const uint32_t (&r)[16], const uint32_t (&d)[16], const uint4* bu, float c2, float lse2, float dls, float scale,
                                         uint32_t km, bool causal, int col0, int i, float (&out_s)[16], const uint2 dkey, const TcDrop& drop, int Tk) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = fmaf(__uint_as_float(r[j]), c2, -lse2);
    if (HAS_BIAS) x += bias_at(bu, j);
    float pe = fast_ex2(x);
    float dp = __uint_as_float(d[j]);
    if (DROP) {  // dP = keep / (1 - p) * dP_drop
      const uint32_t ij = (uint32_t)i * (uint32_t)Tk + (uint32_t)(col0 + j);
      dp = tc_keep(dkey, ij, drop.thresh32) ? dp * drop.inv_keep : 0.f;
    }
    float v = pe * fmaf(dp, scale, -dls);  // scale * P * (dP - delta)
    if (MASKED) {  // (columns past the block's keys hold stale TMEM data: select, never multiply)
      bool ok = (km >> j) & 1u;
      if (causal) ok = ok && (col0 + j <= i);
      v = ok ? v : 0.f;
    }
    out_s[j] = v;
  }
}

// ===================================================================================== backward: dQ
// rows = 128 queries of the tile; loop over 64-key blocks.
// TMEM columns: S / dS [0, 64) | dP [64, 128) | dQ [128, 192)
template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(kThreads, 2)
attn_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const TcParams p) {
  constexpr int BM = 128, BN = 64, STAGES = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // 128 x 128 B
  uint8_t* sDO = sQ + BM * 128;                    // 128 x 128 B
  uint8_t* sK = sDO + BM * 128;                    // STAGES x (64 x 128 B)
  uint8_t* sV = sK + STAGES * BN * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * BN * 128);
  uint64_t* q_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + STAGES;
  uint64_t* sdp_full = empty + STAGES;             // S and dP of the block are in TMEM
  uint64_t* ds_ready = sdp_full + 1;               // dS (bf16) is in TMEM (4 warp arrivals)
  uint64_t* dq_full = ds_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dq_full + 1);
  float* cs_s = reinterpret_cast<float*>(tmem_ptr + 2);  // [4][64] column-sum staging
  uint32_t* kmask_s = reinterpret_cast<uint32_t*>(cs_s + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BM;
  int nkb = (p.Tk + BN - 1) / BN;
  if (p.causal) nkb = min(nkb, (q0 + BM - 1) / BN + 1);
  const int kwords = nkb * (BN / 32);

  pdl_launch();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 4);
    mbar_init(dq_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();
  build_kmask(p, b, kmask_s, kwords);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_dQ = tmem_base + 128;

  if (warp == 0) {
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(q_full, 2 * BM * 128);
      tma_load_3d(sQ, &map_q, q_full, h * 64, q0, b);
      tma_load_3d(sDO, &map_do, q_full, h * 64, q0, b);
    }
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(empty + st, ph ^ 1);
      if (issuer) {
        mbar_expect_tx(full + st, 2 * BN * 128);
        tma_load_3d(sK + st * BN * 128, &map_k, full + st, h * 64, kb * BN, b);
        tma_load_3d(sV + st * BN * 128, &map_v, full + st, h * 64, kb * BN, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool issuer = elect_one();
    mbar_wait(q_full, 0);
    tc_fence_after();
    const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO);
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      const int n = min(BN, ((p.Tk - kb * BN) + 15) & ~15);
      mbar_wait(full + st, ph);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(sK + st * BN * 128), v_addr = smem_u32(sV + st * BN * 128);
      if (issuer) {
        const uint32_t idesc = make_idesc(n, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_S, sdesc(q_addr + k * 32), sdesc(k_addr + k * 32), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_dP, sdesc(do_addr + k * 32), sdesc(v_addr + k * 32), idesc, k != 0);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(ds_ready, kb & 1);
      tc_fence_after();
      if (issuer) {
        const uint32_t idesc = make_idesc(64, true);  // dQ += dS K: B = the K block read MN-major (N = head dim)
        for (int ks = 0; ks < n / 16; ++ks) umma_ts(tmem_dQ, tmem_S + ks * 8, sdesc(k_addr + ks * 2048, BN * 128), idesc, (kb | ks) != 0);
        umma_commit(empty + st);
        if (kb == nkb - 1) umma_commit(dq_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int i = q0 + row;
    const bool row_ok = i < p.Tq;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float c2 = p.scale * kLog2e;
    const __half* brow = HAS_BIAS ? p.bias + (int64_t)h * p.bias_hs + (int64_t)(row_ok ? i : 0) * p.bias_ld : nullptr;
    uint2 dkey = make_uint2(0u, 1u);
    if (DROP) dkey = tc_drop_key(p.drop, b * p.H + h);
    // delta_i = sum_d dO[i, d] * O[i, d]; lse in the log2 domain
    float dl = 0.f, lse2 = 1e30f;  // rows without any visible key (or past Tq): every probability becomes ex2(-huge) = 0
    if (row_ok) {
      const bf16* dop = p.d_o + (int64_t)b * p.do_bs + (int64_t)i * p.do_rs + h * 64;
      const bf16* op = p.o + (int64_t)b * p.o_bs + (int64_t)i * p.o_rs + h * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        const f8 x = load8(dop + c), y = load8(op + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) dl = fmaf(x.v[j], y.v[j], dl);
      }
      const float l = p.lse[((int64_t)b * p.H + h) * p.Tq + i];
      if (l != -INFINITY) lse2 = l * kLog2e;
      p.delta[((int64_t)b * p.H + h) * p.Tq + i] = dl;
    }
    bf16* dsrow = (HAS_BIAS && p.ds != nullptr && row_ok) ? p.ds + (((int64_t)b * p.H + h) * p.Tq + i) * p.bias_ld : nullptr;
    const float dls = dl * p.scale, inv_scale = 1.0f / p.scale;
    for (int kb = 0; kb < nkb; ++kb) {
      const int k0 = kb * BN;
      const int n = min(BN, ((p.Tk - k0) + 15) & ~15);
      const int nch = n >> 4;  // 16-column chunks, software-pipelined: chunk ch + 1 is in flight while chunk ch is processed
      uint4 bu[2][HAS_BIAS ? 2 : 1];
      if constexpr (HAS_BIAS) load_bias_raw<2>(brow, k0, p.bias_ld, row_ok, bu[0]);
      mbar_wait(sdp_full, kb & 1);
      tc_fence_after();
      uint32_t r[2][16], d[2][16];
      tmem_ld16_nowait(tmem_S + lane_addr, r[0]);
      tmem_ld16_nowait(tmem_dP + lane_addr, d[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < nch) {
          const int c0 = ch * 16;
          uint32_t pk[8];
          tmem_ld_wait();
          if (ch + 1 < nch) {
            tmem_ld16_nowait(tmem_S + lane_addr + c0 + 16, r[(ch + 1) & 1]);
            tmem_ld16_nowait(tmem_dP + lane_addr + c0 + 16, d[(ch + 1) & 1]);
            if constexpr (HAS_BIAS) load_bias_raw<2>(brow, k0 + c0 + 16, p.bias_ld, row_ok, bu[(ch + 1) & 1]);
          }
          const uint32_t km = (kmask_s[(k0 + c0) >> 5] >> (c0 & 16)) & 0xffffu;
          const bool diag = p.causal && (k0 + c0 + 15 > q0);
          float dsv[16];  // scale * dS
          if (diag || km != 0xffffu) dq_chunk<true, HAS_BIAS, DROP>(r[ch & 1], d[ch & 1], bu[ch & 1], c2, lse2, dls, p.scale, km, p.causal, k0 + c0, i, dsv, dkey, p.drop, p.Tk);
          else dq_chunk<false, HAS_BIAS, DROP>(r[ch & 1], d[ch & 1], bu[ch & 1], c2, lse2, dls, p.scale, km, p.causal, k0 + c0, i, dsv, dkey, p.drop, p.Tk);
          if (HAS_BIAS && dsrow != nullptr) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              if (k0 + c0 + 8 * v < p.bias_ld) {
                f8 o8;
#pragma unroll
                for (int e = 0; e < 8; ++e) o8.v[e] = dsv[8 * v + e] * inv_scale;
                store8(dsrow + k0 + c0 + 8 * v, o8);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_bf16(dsv[j], dsv[j + 1]);
          tmem_st8(tmem_S + lane_addr + (c0 >> 1), pk);  // columns [8 ch, 8 ch + 8): below every chunk still to be read
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    float lo[32], hi[32];
    {
      uint32_t r0[32], r1[32];
      tmem_ld32_nowait(tmem_dQ + lane_addr, r0);
      tmem_ld32(tmem_dQ + lane_addr + 32, r1);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        lo[j] = __uint_as_float(r0[j]);
        hi[j] = __uint_as_float(r1[j]);
      }
    }
    if (row_ok) {
      bf16* dqp = p.dq + (int64_t)b * p.dq_bs + (int64_t)i * p.dq_rs + h * 64;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        f8 v, w;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v.v[e] = lo[j + e];
          w.v[e] = hi[j + e];
        }
        store8(dqp + j, v);
        store8(dqp + 32 + j, w);
      }
    }
    if (p.dq_colsum != nullptr)  // rows past Tq hold exact zeros (zero-filled Q / dO rows)
      tile_colsum_out(lo, hi, cs_s, p.dq_colsum + ((int64_t)b * gridDim.x + qt) * (p.H * 64) + h * 64);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// One 16-column chunk of the dK/dV kernel on transposed scores: rows = keys, columns = queries (their lse * log2(e) and
// scale * delta come from shared memory, broadcast reads).  out_p: P_drop^T (for dV), out_s: scale * dS^T (for dK).
template <bool MASKED, bool HAS_BIAS, bool DROP>
__device__ __forceinline__ void dkv_chunk(const uint32_t (&r)[16], const uint32_t (&d)[16], const uint4* bu, float c2, const float* __restrict__ lse2,
                                          const float* __restrict__ dls, float scale, bool key_ok, bool causal, int qcol0, int jkey, int Tq,
                                          float (&out_p)[16], float (&out_s)[16], const uint2 dkey, const TcDrop& drop, int Tk) {
#pragma unroll
  for (int e4 = 0; e4 < 16; e4 += 4) {
    const float4 l4 = *reinterpret_cast<const float4*>(lse2 + e4);
    const float4 d4 = *reinterpret_cast<const float4*>(dls + e4);
    const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e4 + u;
      float x = fmaf(__uint_as_float(r[e]), c2, -lq[u]);
      if (HAS_BIAS) x += bias_at(bu, e);
      const float pe = fast_ex2(x);
      float dp = __uint_as_float(d[e]);
      float pv = pe;
      if (DROP) {
        const uint32_t ij = (uint32_t)(qcol0 + e) * (uint32_t)Tk + (uint32_t)jkey;
        const bool keep = tc_keep(dkey, ij, drop.thresh32);
        pv = keep ? pe * drop.inv_keep : 0.f;  // P_drop (dV = P_drop^T dO)
        dp = keep ? dp * drop.inv_keep : 0.f;  // dP = keep / (1 - p) * dP_drop
      }
      float sv = pe * fmaf(dp, scale, -dq4[u]);  // scale * P * (dP - delta)
      if (MASKED) {  // (stale TMEM columns: select, never multiply)
        bool ok = key_ok && (qcol0 + e < Tq);
        if (causal) ok = ok && (jkey <= qcol0 + e);
        pv = ok ? pv : 0.f;
        sv = ok ? sv : 0.f;
      }
      out_p[e] = pv;
      out_s[e] = sv;
    }
  }
}

// ===================================================================================== backward: dK / dV
// rows = 128 keys of the tile; loop over 64-query blocks on TRANSPOSED scores S^T[key, query].
// TMEM columns: S^T / P^T [0, 64) | dP^T / dS^T [64, 128) | dK [128, 192) | dV [192, 256)
template <bool HAS_BIAS, bool DROP>
__global__ void __launch_bounds__(kThreads, 2)
attn_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                       const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do, const TcParams p) {
  constexpr int BM = 128, BQ = 64, STAGES = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                              // 128 x 128 B
  uint8_t* sV = sK + BM * 128;
  uint8_t* sQ = sV + BM * 128;                     // STAGES x (64 x 128 B)
  uint8_t* sDO = sQ + STAGES * BQ * 128;
  float* lse_s = reinterpret_cast<float*>(sDO + STAGES * BQ * 128);  // [STAGES][64] lse * log2(e) of the block's queries
  float* dl_s = lse_s + STAGES * BQ;                                  // [STAGES][64] scale * delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(dl_s + STAGES * BQ);
  uint64_t* kv_full = bars;
  uint64_t* full = bars + 1;
  uint64_t* empty = full + STAGES;
  uint64_t* sdp_full = empty + STAGES;
  uint64_t* ds_ready = sdp_full + 1;
  uint64_t* acc_full = ds_ready + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* cs_s = reinterpret_cast<float*>(tmem_ptr + 2);  // [4][64] column-sum staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * BM;
  const int nqb = (p.Tq + BQ - 1) / BQ;
  const int qb0 = p.causal ? min(nqb, k0 / BQ) : 0;  // queries before the tile's first key never see these keys

  pdl_launch();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    mbar_init(kv_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full + i, 2);  // the TMA transaction arrival + the producer warp's lse / delta rows
      mbar_init(empty + i, 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 4);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 64, tmem_dK = tmem_base + 128, tmem_dV = tmem_base + 192;

  if (warp == 0) {
    const bool issuer = elect_one();
    if (issuer) {
      mbar_expect_tx(kv_full, 2 * BM * 128);
      tma_load_3d(sK, &map_k, kv_full, h * 64, k0, b);
      tma_load_3d(sV, &map_v, kv_full, h * 64, k0, b);
    }
    for (int qb = qb0; qb < nqb; ++qb) {
      const int it = qb - qb0;
      const int st = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(empty + st, ph ^ 1);
      if (issuer) {
        mbar_expect_tx(full + st, 2 * BQ * 128);
        tma_load_3d(sQ + st * BQ * 128, &map_q, full + st, h * 64, qb * BQ, b);
        tma_load_3d(sDO + st * BQ * 128, &map_do, full + st, h * 64, qb * BQ, b);
      }
      // per-query statistics of the block (the dQ kernel wrote delta before this kernel started)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int iq = qb * BQ + t * 32 + lane;
        float l2 = 1e30f, dlv = 0.f;
        if (iq < p.Tq) {
          const float l = p.lse[((int64_t)b * p.H + h) * p.Tq + iq];
          if (l != -INFINITY) l2 = l * kLog2e;
          dlv = p.delta[((int64_t)b * p.H + h) * p.Tq + iq] * p.scale;
        }
        lse_s[st * BQ + t * 32 + lane] = l2;
        dl_s[st * BQ + t * 32 + lane] = dlv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full + st);
    }
  } else if (warp == 1) {
    const bool issuer = elect_one();
    mbar_wait(kv_full, 0);
    tc_fence_after();
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    for (int qb = qb0; qb < nqb; ++qb) {
      const int it = qb - qb0;
      const int st = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      const int n = min(BQ, ((p.Tq - qb * BQ) + 15) & ~15);
      mbar_wait(full + st, ph);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ + st * BQ * 128), do_addr = smem_u32(sDO + st * BQ * 128);
      if (issuer) {
        const uint32_t idesc = make_idesc(n, false);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_S, sdesc(k_addr + k * 32), sdesc(q_addr + k * 32), idesc, k != 0);    // S^T = K Q^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_dP, sdesc(v_addr + k * 32), sdesc(do_addr + k * 32), idesc, k != 0);  // dP^T = V dO^T
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(ds_ready, it & 1);
      tc_fence_after();
      if (issuer) {
        const uint32_t idesc = make_idesc(64, true);
        for (int ks = 0; ks < n / 16; ++ks) umma_ts(tmem_dV, tmem_S + ks * 8, sdesc(do_addr + ks * 2048, BQ * 128), idesc, (it | ks) != 0);  // dV += P^T dO
        for (int ks = 0; ks < n / 16; ++ks) umma_ts(tmem_dK, tmem_dP + ks * 8, sdesc(q_addr + ks * 2048, BQ * 128), idesc, (it | ks) != 0);  // dK += dS^T Q
        umma_commit(empty + st);
        if (qb == nqb - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int j = k0 + row;  // this thread's key
    bool key_ok = j < p.Tk;
    if (key_ok && p.kpm != nullptr) key_ok = p.kpm[(int64_t)b * p.Tk + j] == 0;
    const bool row_ok = j < p.Tk;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float c2 = p.scale * kLog2e;
    const __half* brow = HAS_BIAS ? p.bias_t + (int64_t)h * p.bias_t_hs + (int64_t)(row_ok ? j : 0) * p.bias_t_ld : nullptr;
    uint2 dkey = make_uint2(0u, 1u);
    if (DROP) dkey = tc_drop_key(p.drop, b * p.H + h);
    const bool warp_keys_ok = __all_sync(0xffffffffu, key_ok);
    for (int qb = qb0; qb < nqb; ++qb) {
      const int it = qb - qb0;
      const int st = it % STAGES;
      const int q0 = qb * BQ;
      const int n = min(BQ, ((p.Tq - q0) + 15) & ~15);
      const int nch = n >> 4;
      uint4 bu[2][HAS_BIAS ? 2 : 1];
      if constexpr (HAS_BIAS) load_bias_raw<2>(brow, q0, p.bias_t_ld, row_ok, bu[0]);
      mbar_wait(sdp_full, it & 1);  // (the MMA issuer waited for full[st]: the lse / delta rows of this stage are visible)
      tc_fence_after();
      const float* lse2 = lse_s + st * BQ;
      const float* dls = dl_s + st * BQ;
      uint32_t r[2][16], d[2][16];
      tmem_ld16_nowait(tmem_S + lane_addr, r[0]);
      tmem_ld16_nowait(tmem_dP + lane_addr, d[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < nch) {
          const int c0 = ch * 16;
          uint32_t pp[8], pd[8];
          tmem_ld_wait();
          if (ch + 1 < nch) {
            tmem_ld16_nowait(tmem_S + lane_addr + c0 + 16, r[(ch + 1) & 1]);
            tmem_ld16_nowait(tmem_dP + lane_addr + c0 + 16, d[(ch + 1) & 1]);
            if constexpr (HAS_BIAS) load_bias_raw<2>(brow, q0 + c0 + 16, p.bias_t_ld, row_ok, bu[(ch + 1) & 1]);
          }
          float pv[16], dsv[16];
          // fast path: every key row of this warp valid, every query column of the chunk < Tq, no causal diagonal
          const bool masked = !warp_keys_ok || (q0 + c0 + 16 > p.Tq) || (p.causal && (k0 + BM - 1 > q0 + c0));
          if (masked) dkv_chunk<true, HAS_BIAS, DROP>(r[ch & 1], d[ch & 1], bu[ch & 1], c2, lse2 + c0, dls + c0, p.scale, key_ok, p.causal, q0 + c0, j, p.Tq, pv, dsv, dkey, p.drop, p.Tk);
          else dkv_chunk<false, HAS_BIAS, DROP>(r[ch & 1], d[ch & 1], bu[ch & 1], c2, lse2 + c0, dls + c0, p.scale, key_ok, p.causal, q0 + c0, j, p.Tq, pv, dsv, dkey, p.drop, p.Tk);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            pp[e >> 1] = pack_bf16(pv[e], pv[e + 1]);
            pd[e >> 1] = pack_bf16(dsv[e], dsv[e + 1]);
          }
          tmem_st8(tmem_S + lane_addr + (c0 >> 1), pp);
          tmem_st8(tmem_dP + lane_addr + (c0 >> 1), pd);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    if (qb0 < nqb) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* base = which == 0 ? p.dk + (int64_t)b * p.dk_bs + (int64_t)j * p.dk_rs : p.dv + (int64_t)b * p.dv_bs + (int64_t)j * p.dv_rs;
      float lo[32], hi[32];
      if (qb0 < nqb) {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait((which == 0 ? tmem_dK : tmem_dV) + lane_addr, r0);
        tmem_ld32((which == 0 ? tmem_dK : tmem_dV) + lane_addr + 32, r1);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          lo[e] = __uint_as_float(r0[e]);
          hi[e] = __uint_as_float(r1[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) lo[e] = hi[e] = 0.f;
      }
      if (row_ok) {
        bf16* gp = base + h * 64;
#pragma unroll
        for (int e0 = 0; e0 < 32; e0 += 8) {
          f8 v, w;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            v.v[e] = lo[e0 + e];
            w.v[e] = hi[e0 + e];
          }
          store8(gp + e0, v);
          store8(gp + 32 + e0, w);
        }
      }
      float* csout = which == 0 ? p.dk_colsum : p.dv_colsum;
      if (csout != nullptr)  // rows past Tk hold exact zeros
        tile_colsum_out(lo, hi, cs_s, csout + ((int64_t)b * gridDim.x + kt) * (p.H * 64) + h * 64);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

// ===================================================================================== position-bias tiles
// bias[h, i, j] = abs_scale * abs[h, i, j] + table[idx[i, j], h]   (fp16; columns [Tk, ld) zero) and the transposed tile
// bias_t[h, j, i] for the dK/dV kernel.  abs: fp32 [H, Tq, Tk] (abs-pos term pq . pk per head, general.py:223-243) or NULL;
// idx: int32 [Tq, Tk] bucket ids (-1 = none) or NULL with table fp32 / bf16 [n_buckets, H] (general.py:270-280).
template <typename TT>
__global__ void attn_bias_build_kernel(const float* __restrict__ abs, int abs_ld, float abs_scale, const int* __restrict__ idx, const TT* __restrict__ table,
                                       int H, int Tq, int Tk, __half* __restrict__ out, int ld, __half* __restrict__ out_t, int ld_t) {
  __shared__ float tile[32][33];
  pdl_launch();
  pdl_wait();
  const int h = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    float v = 0.f;
    if (i < Tq && j < Tk) {
      if (abs != nullptr) v = abs[((int64_t)h * Tq + i) * abs_ld + j] * abs_scale;
      if (idx != nullptr) {
        const int id = idx[(int64_t)i * Tk + j];
        if (id >= 0) v += (float)table[(int64_t)id * H + h];
      }
    }
    v *= kLog2e;  // the attention kernels work in the log2 domain: p = ex2(x)
    tile[r][tx] = v;
    if (i < Tq && j < ld) out[((int64_t)h * Tq + i) * ld + j] = __float2half_rn(v);
  }
  __syncthreads();
  if (out_t != nullptr) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int j = j0 + r, i = i0 + tx;
      if (j < Tk && i < ld_t) out_t[((int64_t)h * Tk + j) * ld_t + i] = __float2half_rn(i < Tq ? tile[tx][r] : 0.f);
    }
  }
}

// g[h, i, j] = sum_b ds[b, h, i, j];  dabs[h, i, j] (+)= abs_scale * g;  dtable[idx[i, j], h] += g  (fp32 atomics)
__global__ void attn_bias_bwd_kernel(const bf16* __restrict__ ds, int B, int H, int Tq, int Tk, int ld, const int* __restrict__ idx,
                                     float* __restrict__ dtable, float* __restrict__ dabs, int abs_ld, float abs_scale, int dabs_accum) {
  pdl_launch();
  pdl_wait();
  const int64_t per_b = (int64_t)H * Tq * ld;
  const int64_t total = (int64_t)H * Tq * (ld / 8);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int jc = (int)(t % (ld / 8)) * 8;
    const int64_t hi = t / (ld / 8);
    const int i = (int)(hi % Tq), h = (int)(hi / Tq);
    float g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = 0.f;
    for (int b = 0; b < B; ++b) {
      const f8 v = load8(ds + b * per_b + hi * ld + jc);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] += v.v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = jc + e;
      if (j >= Tk) continue;
      if (dabs != nullptr) {
        float* dp = dabs + ((int64_t)h * Tq + i) * abs_ld + j;
        *dp = dabs_accum ? *dp + abs_scale * g[e] : abs_scale * g[e];
      }
      if (idx != nullptr && g[e] != 0.f) {
        const int id = idx[(int64_t)i * Tk + j];
        if (id >= 0) atomicAdd(dtable + (int64_t)id * H + h, g[e]);
      }
    }
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {  // cuTensorMapEncodeTiled needs a current context on the calling thread (see gemm.cu)
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

// bf16 [B, T, cols] with element strides (bs, rs, 1); box = 64 columns x `rows` rows x 1 batch, SWIZZLE_128B.
// Returns false when the driver refuses the shape / strides (the caller falls back to the mma.sync kernels).
bool make_map3(CUtensorMap* map, const void* ptr, int cols, int T, int B, int64_t rs, int64_t bs, int rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (((uintptr_t)ptr & 15) != 0 || rs % 8 != 0 || bs % 8 != 0 || rs <= 0) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)rs * 2, (cuuint64_t)(bs > 0 ? bs : rs * T) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <typename K>
int tc_launch(K kern, dim3 grid, int smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const TcParams& p) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_attn (tcgen05): cudaFuncSetAttribute");
  e = ofab_launch(kern, grid, dim3(kThreads), (size_t)smem, st, a, b, c, p);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_attn (tcgen05) launch");
  return OFAB_OK;
}
template <typename K>
int tc_launch4(K kern, dim3 grid, int smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& d,
               const TcParams& p) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_attn (tcgen05): cudaFuncSetAttribute");
  e = ofab_launch(kern, grid, dim3(kThreads), (size_t)smem, st, a, b, c, d, p);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_attn (tcgen05) launch");
  return OFAB_OK;
}

bool fill_tc(const ofab_attn_fwd_args* a, TcParams& p, bool& drop_on) {
  p = TcParams{};
  p.B = a->B; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk;
  p.scale = a->scale;
  p.kpm = a->kpm;
  p.causal = a->causal;
  p.bias = (const __half*)a->bias;
  p.bias_hs = a->bias_hs;
  p.bias_ld = a->bias_ld;
  p.bias_t = (const __half*)a->bias_t;
  p.bias_t_hs = a->bias_t_hs;
  p.bias_t_ld = a->bias_t_ld;
  p.o = (bf16*)a->o; p.o_bs = a->o_bs; p.o_rs = a->o_rs;
  p.lse = a->lse;
  drop_on = a->drop != nullptr && a->drop->p > 0.f;
  p.drop = TcDrop{nullptr, 0u, 0u, 1.0f};
  if (drop_on) {
    p.drop.state = reinterpret_cast<const unsigned long long*>(a->drop->state);
    p.drop.site = a->drop->site;
    const double t = (double)a->drop->p * 4294967296.0;
    p.drop.thresh32 = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
    p.drop.inv_keep = 1.0f / (1.0f - a->drop->p);
  }
  return true;
}

}  // namespace

// (declared in common.cuh; called by the dispatchers in attn.cu)
// Returns 1 when the tcgen05 kernels can take this problem (no structured position terms; strides a 3-D tensor map accepts).
int ofab_attn_tc_eligible(const ofab_attn_fwd_args* a) {
  if (a->pq != nullptr || a->rp_idx != nullptr) return 0;
  if (a->bias != nullptr && (a->bias_ld % 8 != 0 || a->bias_ld < a->Tk || ((uintptr_t)a->bias & 15) != 0 || a->bias_hs % 8 != 0)) return 0;
  static int legacy = -1;
  if (legacy < 0) {
    const char* e = getenv("OFAB_ATTN_LEGACY");
    legacy = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  return legacy ? 0 : 1;
}

// rc: OFAB_OK, an error, or 1 = "not taken" (tensor maps refused): the caller runs the mma.sync kernels instead.
int ofab_attn_tc_fwd(const ofab_attn_fwd_args* a, ofab_stream_t stream) {
  TcParams p;
  bool drop_on;
  fill_tc(a, p, drop_on);
  CUtensorMap mq, mk, mv;
  if (!make_map3(&mq, a->q, a->H * 64, a->Tq, a->B, a->q_rs, a->q_bs, 128)) return 1;
  if (!make_map3(&mk, a->k, a->H * 64, a->Tk, a->B, a->k_rs, a->k_bs, 64)) return 1;  // 64-key blocks
  if (!make_map3(&mv, a->v, a->H * 64, a->Tk, a->B, a->v_rs, a->v_bs, 64)) return 1;
  const int nkb = (a->Tk + 63) / 64;
  const int smem = 1024 + 2 * 128 * 128 + 2 * 4 * 64 * 128 + 24 * 8 + 16 + (nkb + 1) * 2 * 4;
  // heads per CTA: minimise waves x (set-up + heads x work), set-up being about two heads' worth of work (ncu, profiles/README)
  const int qtiles = (a->Tq + 127) / 128;
  int hpc = 1;
  if (const char* ov = getenv("OFAB_ATTN_HPC")) hpc = atoi(ov) > 0 ? atoi(ov) : 1;
  else {
    const int64_t slots = 2 * (int64_t)ofab_sm_count();
    int64_t best = INT64_MAX;
    for (int c = 1; c <= 6 && c <= a->H; ++c) {
      const int64_t ctas = (int64_t)qtiles * ((a->H + c - 1) / c) * a->B;
      const int64_t cost = ((ctas + slots - 1) / slots) * (2 + c);
      if (cost < best) best = cost, hpc = c;
    }
  }
  p.hpc = hpc;
  dim3 grid(qtiles, (a->H + hpc - 1) / hpc, a->B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool hb = a->bias != nullptr;
  if (hb && drop_on) return tc_launch(attn_tc_fwd_kernel<true, true>, grid, smem, st, mq, mk, mv, p);
  if (hb) return tc_launch(attn_tc_fwd_kernel<true, false>, grid, smem, st, mq, mk, mv, p);
  if (drop_on) return tc_launch(attn_tc_fwd_kernel<false, true>, grid, smem, st, mq, mk, mv, p);
  return tc_launch(attn_tc_fwd_kernel<false, false>, grid, smem, st, mq, mk, mv, p);
}

int ofab_attn_tc_bwd(const ofab_attn_bwd_args* a, ofab_stream_t stream) {
  TcParams p;
  bool drop_on;
  fill_tc(&a->f, p, drop_on);
  p.d_o = (const bf16*)a->d_o; p.do_bs = a->do_bs; p.do_rs = a->do_rs;
  p.delta = a->delta;
  p.dq = (bf16*)a->dq; p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv;
  p.dq_bs = a->dq_bs; p.dq_rs = a->dq_rs; p.dk_bs = a->dk_bs; p.dk_rs = a->dk_rs; p.dv_bs = a->dv_bs; p.dv_rs = a->dv_rs;
  p.ds = (bf16*)a->ds;
  p.dq_colsum = a->dq_colsum; p.dk_colsum = a->dk_colsum; p.dv_colsum = a->dv_colsum;
  const ofab_attn_fwd_args& f = a->f;
  CUtensorMap q128, do128, k64, v64, k128, v128, q64, do64;
  const int C = f.H * 64;
  if (!make_map3(&q128, f.q, C, f.Tq, f.B, f.q_rs, f.q_bs, 128) || !make_map3(&do128, a->d_o, C, f.Tq, f.B, a->do_rs, a->do_bs, 128) ||
      !make_map3(&k64, f.k, C, f.Tk, f.B, f.k_rs, f.k_bs, 64) || !make_map3(&v64, f.v, C, f.Tk, f.B, f.v_rs, f.v_bs, 64) ||
      !make_map3(&k128, f.k, C, f.Tk, f.B, f.k_rs, f.k_bs, 128) || !make_map3(&v128, f.v, C, f.Tk, f.B, f.v_rs, f.v_bs, 128) ||
      !make_map3(&q64, f.q, C, f.Tq, f.B, f.q_rs, f.q_bs, 64) || !make_map3(&do64, a->d_o, C, f.Tq, f.B, a->do_rs, a->do_bs, 64))
    return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool hb = f.bias != nullptr;
  OFAB_REQUIRE(!hb || f.bias_t != nullptr, "ofab_attn_bwd: bias_t (the transposed bias tile) is required with bias");
  OFAB_REQUIRE(!hb || (f.bias_t_ld % 8 == 0 && f.bias_t_ld >= f.Tq && ((uintptr_t)f.bias_t & 15) == 0 && f.bias_t_hs % 8 == 0),
               "ofab_attn_bwd: bias_t needs 16-byte aligned rows with bias_t_ld >= Tq");
  const int nkb = (f.Tk + 63) / 64;
  const int smem_q = 1024 + 2 * 128 * 128 + 2 * 2 * 64 * 128 + 16 * 8 + 16 + 1024 + nkb * 2 * 4;
  const int smem_kv = 1024 + 2 * 128 * 128 + 2 * 2 * 64 * 128 + 2 * 2 * 64 * 4 + 16 * 8 + 16 + 1024;
  dim3 gq((f.Tq + 127) / 128, f.H, f.B), gkv((f.Tk + 127) / 128, f.H, f.B);
  int rc;
#define TC_BWD(HB, D)                                                                                      \
  {                                                                                                        \
    if ((rc = tc_launch4(attn_tc_bwd_dq_kernel<HB, D>, gq, smem_q, st, q128, do128, k64, v64, p))) return rc; \
    return tc_launch4(attn_tc_bwd_dkv_kernel<HB, D>, gkv, smem_kv, st, k128, v128, q64, do64, p);          \
  }
  if (hb && drop_on) TC_BWD(true, true)
  if (hb) TC_BWD(true, false)
  if (drop_on) TC_BWD(false, true)
  TC_BWD(false, false)
#undef TC_BWD
}

extern "C" int ofab_attn_bias_build(const float* abs, int abs_ld, float abs_scale, const int32_t* idx, const void* table, int table_dt, int n_buckets, int H,
                                    int Tq, int Tk, void* out, int ld, void* out_t, int ld_t, ofab_stream_t stream) {
  OFAB_REQUIRE(H > 0 && Tq > 0 && Tk > 0 && out != nullptr, "ofab_attn_bias_build: bad arguments");
  OFAB_REQUIRE((idx == nullptr) == (table == nullptr), "ofab_attn_bias_build: idx and table go together");
  OFAB_REQUIRE(abs == nullptr || abs_ld >= Tk, "ofab_attn_bias_build: abs_ld=%d < Tk", abs_ld);
  OFAB_REQUIRE(ld >= Tk && ld % 8 == 0 && (out_t == nullptr || (ld_t >= Tq && ld_t % 8 == 0)), "ofab_attn_bias_build: ld=%d / ld_t=%d must be multiples of 8 covering the row", ld, ld_t);
  (void)n_buckets;
  dim3 grid((ld + 31) / 32, (Tq + 31) / 32, H);
  // the transposed tile pads its rows (queries) to ld_t: make the grid cover max(Tq, ld_t) query columns
  if (out_t != nullptr && (ld_t + 31) / 32 > (int)grid.y) grid.y = (ld_t + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
  if (table_dt == OFAB_BF16)
    ofab_launch(attn_bias_build_kernel<bf16>, grid, dim3(256), 0, st, abs, abs_ld, abs_scale, idx, (const bf16*)table, H, Tq, Tk, (__half*)out, ld, (__half*)out_t, ld_t);
  else
    ofab_launch(attn_bias_build_kernel<float>, grid, dim3(256), 0, st, abs, abs_ld, abs_scale, idx, (const float*)table, H, Tq, Tk, (__half*)out, ld, (__half*)out_t, ld_t);
  OFAB_LAUNCH_CHECK("ofab_attn_bias_build");
  return OFAB_OK;
}

extern "C" int ofab_attn_bias_bwd(const void* ds, int B, int H, int Tq, int Tk, int ld, const int32_t* idx, float* dtable, float* dabs,
                                  int abs_ld, float abs_scale, int dabs_accum, ofab_stream_t stream) {
  OFAB_REQUIRE(ds != nullptr && B > 0 && H > 0 && Tq > 0 && Tk > 0 && ld >= Tk && ld % 8 == 0, "ofab_attn_bias_bwd: bad arguments");
  OFAB_REQUIRE((idx == nullptr) == (dtable == nullptr), "ofab_attn_bias_bwd: idx and dtable go together");
  OFAB_REQUIRE(dabs == nullptr || abs_ld >= Tk, "ofab_attn_bias_bwd: abs_ld=%d < Tk", abs_ld);
  const int64_t total = (int64_t)H * Tq * (ld / 8);
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)ofab_sm_count() * 8 ? want : (int64_t)ofab_sm_count() * 8);
  ofab_launch(attn_bias_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)ds, B, H, Tq, Tk, ld, idx, dtable, dabs, abs_ld, abs_scale, dabs_accum);
  OFAB_LAUNCH_CHECK("ofab_attn_bias_bwd");
  return OFAB_OK;
}
