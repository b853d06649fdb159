// Fused attention forward / backward (flash-style) for head_dim 64.
//
// Scores, position biases and probabilities live only in registers; HBM traffic is q/k/v/o (+ the
// int32 bucket-id map shared by every head and layer).  The absolute-position term of OFA
// (general.py:223-243: (W_q pos * s)(W_k pos)^T) rides along as 64 extra contraction columns of the
// QK^T product (pq/pk), the relative-position term is a per-head table column staged in shared
// memory and indexed by the bucket id of (i, j).
//
// Tensor-core path here is mma.sync m16n8k16 (bf16 in, fp32 accumulate): at OFA's sequence lengths
// attention is ~1-3 % of the step FLOPs, so it is written for correctness and minimal HBM traffic
// first; the tcgen05 budget goes to the GEMMs (gemm.cu).
//
// Tiles: 64 query rows x 64 keys per step, 4 warps (16 query rows each).  smem tiles are 64 rows of
// 128 bytes with a 16-byte-chunk XOR swizzle so ldmatrix is conflict free.
#include "common.cuh"

namespace {

// resident CTAs per SM the register allocator is asked for (no position columns / with): tools/build_attn_variants.sh
// builds alternatives for A/B runs.  Measured in the whole step (B=64): forward 4 / backward 3 is the best of
// {4,3} x {3,2}; grouping all ldmatrix loads of a k step ahead of its MMAs (profiles/r01_attn_variants.txt) cost
// 120-244 B of spills in the backward kernels and +8 % of their in-step time.
#ifndef ATTN_FWD_MINB
#define ATTN_FWD_MINB 4
#endif
#ifndef ATTN_BWD_MINB
#define ATTN_BWD_MINB 3
#endif
constexpr int TILE = 64;
constexpr int TILE_BYTES = TILE * 128;  // 64 rows x 64 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// async copy of rows [row0, row0+64) x 64 columns (bf16) of a strided global matrix into a swizzled tile.
// rows >= nrows are zero-filled.  All 128 threads participate (4 chunks each).
__device__ __forceinline__ void load_tile_async(uint32_t tile, const bf16* __restrict__ g, int64_t row_stride, int row0, int nrows) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * 128;  // 0..511 : 64 rows x 8 chunks
    const int r = idx >> 3, c = idx & 7;
    const int gr = row0 + r;
    const bool ok = gr < nrows;
    const bf16* src = g + (int64_t)(ok ? gr : 0) * row_stride + c * 8;
    const uint32_t dst = tile_addr(tile, r, c);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
  }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// A fragment: rows [r0, r0+16), k columns [kk*16, kk*16+16) of a tile
__device__ __forceinline__ void frag_a(uint32_t tile, int r0, int kk, uint32_t (&a)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm_x4(tile_addr(tile, r0 + (l & 7) + (mi & 1) * 8, kk * 2 + (mi >> 1)), a);
}
// B fragments (two n-tiles) where the tile is stored [n][k]: n rows [n0, n0+16), k columns [kk*16, +16)
__device__ __forceinline__ void frag_b(uint32_t tile, int n0, int kk, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm_x4(tile_addr(tile, n0 + (l & 7) + (mi >> 1) * 8, kk * 2 + (mi & 1)), b);
}
// B fragments (two n-tiles) where the tile is stored [k][n]: k rows [k0, k0+16), n columns [nn*16, +16)
__device__ __forceinline__ void frag_bt(uint32_t tile, int k0, int nn, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm_x4_t(tile_addr(tile, k0 + (l & 7) + (mi & 1) * 8, nn * 2 + (mi >> 1)), b);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnCommon {
  int B, H, Tq, Tk;
  const bf16 *q, *k, *v, *pq, *pk;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, pq_bs, pq_rs, pk_bs, pk_rs;
  const int16_t* rp_idx;    // [Tq, rp_ld] bucket ids, -1 = none (columns >= Tk hold -1; rp_ld even)
  const int16_t* rp_idx_t;  // the same map transposed, [Tk, rp_ld_t] (dK/dV kernel: its accumulator rows are keys)
  int rp_ld, rp_ld_t;
  const float* table;
  int n_buckets;
  const uint8_t* kpm;
  int causal;
  float scale;
};

// ---- attention dropout (multihead_attention.py:335: attn_probs = dropout(softmax(scores))) ------------------------
// The keep bit of probability (b, h, i, j) is a keyed 32-bit hash of i * Tk + j; the two key words come from one
// Philox4x32-7 call per thread on (seed, step, site, b * H + h), so the forward kernel and both backward kernels --
// whose accumulator layouts differ (queries x keys vs keys x queries) -- regenerate the same bits element by element.
// The softmax denominator and the saved log-sum-exp stay those of the undropped scores; backward keeps
// delta = rowsum(dO * O) unchanged because rowsum(P * dP) = rowsum(P_drop * dP_drop).
struct AttnDrop {
  const unsigned long long* state;  // device [2]: seed, step (see DropArgs)
  uint32_t site;
  uint32_t thresh32;  // dropped when hash < thresh32  (= p * 2^32)
  float inv_keep;     // 1 / (1 - p)
};
__device__ __forceinline__ uint2 attn_drop_key(const AttnDrop& d, int bh) {
  const unsigned long long seed = d.state[0], step = d.state[1];
  const uint4 r = philox4x32_7(make_uint4((uint32_t)bh, d.site, (uint32_t)step, 0x6A09E667u),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
  return make_uint2(r.x, r.y | 1u);
}
__device__ __forceinline__ bool attn_keep(const uint2 key, uint32_t ij, uint32_t thresh32) {
  uint32_t x = (ij ^ key.x) * 0x9E3779B1u;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x *= key.y;  // odd
  x ^= x >> 16;
  x *= 0xC2B2AE3Du;
  x ^= x >> 15;
  return x >= thresh32;
}

// Key-validity mask of one 64-key tile (bit jl set <=> key k0+jl is in range and not padding), built
// cooperatively by warps 0 and 1 while the tile's cp.async copies are in flight.
__device__ __forceinline__ void build_kmask(const AttnCommon& p, int b, int k0, uint32_t* dst /* smem [2] */) {
  if (threadIdx.x < 64) {
    const int j = k0 + threadIdx.x;
    bool ok = j < p.Tk;
    if (ok && p.kpm != nullptr) ok = p.kpm[(int64_t)b * p.Tk + j] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) dst[threadIdx.x >> 5] = m;
  }
}

// Key-validity bitmap of the whole key axis, built once per CTA: bit j of word j/32 set <=> key j is in range and
// not padding.  (One pass of coalesced byte loads in the prologue instead of a dependent global load per key tile.)
__device__ __forceinline__ void build_kmask_all(const AttnCommon& p, int b, uint32_t* dst /* smem [ceil(Tk/64)*2] */) {
  const int words = ((p.Tk + 63) >> 6) * 2;
  for (int w0 = (threadIdx.x >> 5); w0 < words; w0 += (blockDim.x >> 5)) {
    const int j = w0 * 32 + (threadIdx.x & 31);
    bool ok = j < p.Tk;
    if (ok && p.kpm != nullptr) ok = p.kpm[(int64_t)b * p.Tk + j] == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0) dst[w0] = m;
  }
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Bucket ids of this thread's 32 scores of a 16 x 64 warp tile as 16 packed pairs (two adjacent columns per 32-bit
// load).  Issued BEFORE the tile's MMAs so the gather latency hides behind them: with one dependent int32 load per score
// inside finish_tile the position-bias kernels spent ~5 of ~11 stall cycles per issued instruction on the long
// scoreboard (profiles/r01_ncu_full_v11_attn_modeA.csv).
template <bool TRANSPOSED>
__device__ __forceinline__ void load_idx_pairs(const AttnCommon& p, int q0, int k0, int row0, uint32_t (&ip)[8][2]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int16_t* base = TRANSPOSED ? p.rp_idx_t : p.rp_idx;
  const int ld = TRANSPOSED ? p.rp_ld_t : p.rp_ld, nrow = TRANSPOSED ? p.Tk : p.Tq;
  const int a0 = (TRANSPOSED ? k0 : q0) + row0 + g, b0 = (TRANSPOSED ? q0 : k0) + 2 * t;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int a = a0 + 8 * hr, b = b0 + nt * 8;
      ip[nt][hr] = (a < nrow && b < ld) ? __ldg(reinterpret_cast<const unsigned int*>(base + (int64_t)a * ld + b)) : 0xFFFFFFFFu;
    }
}

// Score post-processing shared by forward and both backward kernels, on one warp's 16 x 64 accumulator tile.
// Everything downstream works in the LOG2 domain: x2 = log2(e) * (scale * acc + table[bucket(i, j)]), masked -> -inf,
// so that a probability is ONE ffma + ONE ex2:  p = ex2(x * mult - m2).
//   return value `mult`: the factor that takes the tile's values to the log2 domain --
//     fast path (no table, no masked key in the tile, tile not on the causal diagonal): the accumulators are left
//       untouched and mult = scale * log2(e);
//     otherwise the values are rewritten as x2 (bias added, mask applied) and mult = 1.
//   TRANSPOSED = false: accumulator rows are queries (row0 = first query row of the warp), columns keys.
//   TRANSPOSED = true : accumulator rows are keys   (row0 = first key row of the warp),   columns queries.
//   tab_s holds the table column pre-multiplied by log2(e).
//   ip: the tile's bucket ids from load_idx_pairs (HAS_TAB only).
template <bool TRANSPOSED, bool HAS_TAB, typename F>
__device__ __forceinline__ float finish_tile(const AttnCommon& p, float (&s)[8][4], int q0, int k0, int row0, uint64_t kmask,
                                             const float* tab_s, const uint32_t (&ip)[8][2], F&& on_idx) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr bool has_tab = HAS_TAB;
  const float c2 = p.scale * kLog2e;
  const bool diag = p.causal && (k0 + TILE - 1 > q0);  // some (i, j) of this CTA tile may have j > i
  if (!has_tab && !diag && kmask == ~0ull) return c2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = row0 + g + 8 * (e >> 1), c = nt * 8 + 2 * t + (e & 1);
      const int il = TRANSPOSED ? c : r, jl = TRANSPOSED ? r : c;
      const int i = q0 + il, j = k0 + jl;
      bool ok = (kmask >> jl) & 1ull;
      if (p.causal) ok = ok && (j <= i);
      float v = s[nt][e] * c2;
      if (has_tab && ok && i < p.Tq) {
        const int idx = (int)(int16_t)(ip[nt][e >> 1] >> (16 * (e & 1)));
        if (idx >= 0) {
          v += tab_s[idx];
          on_idx(nt, e, idx);
        }
      }
      s[nt][e] = ok ? v : -INFINITY;
    }
  return 1.0f;
}

// ===================================================================================== forward
// smem: Q (NH tiles) | K 2 stages x NH | V 2 stages | key-validity bitmap (Tk bits) | table column (n_buckets floats)
// K/V ring: 2 stages, one __syncthreads per key tile: tile kv+1 is fetched (cp.async) while tile kv is consumed.
template <bool HAS_POS, bool HAS_TAB, bool DROP = false>
__global__ void __launch_bounds__(128, HAS_POS ? ATTN_FWD_MINB - 1 : ATTN_FWD_MINB) attn_fwd_kernel(const AttnCommon p, bf16* __restrict__ o, int64_t o_bs, int64_t o_rs,
                                                                        float* __restrict__ lse, const AttnDrop ad) {
  constexpr int NH = HAS_POS ? 2 : 1;  // 64-wide halves of the QK contraction
  constexpr int NST = 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + NH * TILE_BYTES;
  const uint32_t sV = sK + NST * NH * TILE_BYTES;
  uint32_t* kmask_s = reinterpret_cast<uint32_t*>(smem + ((NST + 1) * NH + NST) * TILE_BYTES);
  float* tab_s = reinterpret_cast<float*>(kmask_s + ((p.Tk + 63) >> 6) * 2);

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qb * TILE;
  const bool warp_live = q0 + warp * 16 < p.Tq;  // tail tile: warps without valid query rows only help loading
  pdl_launch();
  pdl_wait();  // PDL: no global access above

  const bf16* qg = p.q + (int64_t)b * p.q_bs + h * 64;
  const bf16* kg = p.k + (int64_t)b * p.k_bs + h * 64;
  const bf16* vg = p.v + (int64_t)b * p.v_bs + h * 64;
  const bf16* pqg = HAS_POS ? p.pq + (int64_t)b * p.pq_bs + h * 64 : nullptr;
  const bf16* pkg = HAS_POS ? p.pk + (int64_t)b * p.pk_bs + h * 64 : nullptr;

  const int n_kv = p.causal ? min((p.Tk + TILE - 1) / TILE, (q0 + TILE + TILE - 1) / TILE) : (p.Tk + TILE - 1) / TILE;

  auto load_kv = [&](int kvi) {
    const int stg = kvi % NST, k1 = kvi * TILE;
    load_tile_async(sK + (stg * NH) * TILE_BYTES, kg, p.k_rs, k1, p.Tk);
    if (HAS_POS) load_tile_async(sK + (stg * NH + 1) * TILE_BYTES, pkg, p.pk_rs, k1, p.Tk);
    load_tile_async(sV + stg * TILE_BYTES, vg, p.v_rs, k1, p.Tk);
    cp_commit();
  };
  load_tile_async(sQ, qg, p.q_rs, q0, p.Tq);
  if (HAS_POS) load_tile_async(sQ + TILE_BYTES, pqg, p.pq_rs, q0, p.Tq);
  load_kv(0);
  if (HAS_TAB)
    for (int i = threadIdx.x; i < p.n_buckets; i += 128) tab_s[i] = p.table[(int64_t)i * p.H + h] * kLog2e;
  build_kmask_all(p, b, kmask_s);

  float oacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // running max (log2 domain) and sum per query row
  const int row_g[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  uint2 dkey = make_uint2(0u, 1u);
  if (DROP) dkey = attn_drop_key(ad, b * p.H + h);

  for (int kv = 0; kv < n_kv; ++kv) {
    const int st = kv % NST;
    cp_wait<0>();      // tile kv has landed ...
    __syncthreads();   // ... for every thread; and everyone finished tile kv-1 (and, first time, the bitmap / table)
    if (kv + 1 < n_kv) load_kv(kv + 1);  // refill the stage consumed in iteration kv-1

    if (warp_live) {
      const int k0 = kv * TILE;
      const int nk = min(TILE, p.Tk - k0);          // valid keys in this tile
      const int np_n = (nk + 15) >> 4;               // 16-key groups that hold a valid key
      const uint64_t kmask = (uint64_t)kmask_s[2 * kv] | ((uint64_t)kmask_s[2 * kv + 1] << 32);
      uint32_t ip[8][2];
      if (HAS_TAB) load_idx_pairs<false>(p, q0, k0, warp * 16, ip);
      // ---- S = Q K^T (+ PQ PK^T)
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        const uint32_t tq = sQ + hf * TILE_BYTES, tk = sK + (st * NH + hf) * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t a[4];
          frag_a(tq, warp * 16, kk, a);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            if (np < np_n) {
              uint32_t bb[4];
              frag_b(tk, np * 16, kk, bb);
              mma16816(s[2 * np], a, bb[0], bb[1]);
              mma16816(s[2 * np + 1], a, bb[2], bb[3]);
            }
          }
        }
      }
      // ---- bias / mask / online softmax (log2 domain: p = ex2(s * mult - m), one FFMA + one MUFU per score)
      const float mult = finish_tile<false, HAS_TAB>(p, s, q0, k0, warp * 16, kmask, tab_s, ip, [](int, int, int) {});
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      float corr[2], m_use[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r] * mult);  // mult > 0: the max commutes with the scaling
        m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
        corr[r] = fast_ex2(m_run[r] - m_use[r]);  // ex2(-inf) = 0 on the first tile
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
      float ps[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pe = fast_ex2(fmaf(s[nt][e], mult, -m_use[e >> 1]));
          s[nt][e] = pe;
          ps[e >> 1] += pe;
        }
#pragma unroll
      for (int r = 0; r < 2; ++r) l_run[r] += ps[r];
      if (DROP) {  // the row sums above are those of the undropped probabilities; only the P V product sees the mask
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t ij = (uint32_t)row_g[e >> 1] * (uint32_t)p.Tk + (uint32_t)(k0 + nt * 8 + 2 * t + (e & 1));
            s[nt][e] = attn_keep(dkey, ij, ad.thresh32) ? s[nt][e] * ad.inv_keep : 0.f;
          }
      }
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        oacc[dt][0] *= corr[0]; oacc[dt][1] *= corr[0];
        oacc[dt][2] *= corr[1]; oacc[dt][3] *= corr[1];
      }
      // ---- O += P V
      const uint32_t tv = sV + st * TILE_BYTES;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        if (kb < np_n) {
          uint32_t a[4];
          a[0] = pack_bf16(s[2 * kb][0], s[2 * kb][1]);
          a[1] = pack_bf16(s[2 * kb][2], s[2 * kb][3]);
          a[2] = pack_bf16(s[2 * kb + 1][0], s[2 * kb + 1][1]);
          a[3] = pack_bf16(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t bb[4];
            frag_bt(tv, kb * 16, dp, bb);
            mma16816(oacc[2 * dp], a, bb[0], bb[1]);
            mma16816(oacc[2 * dp + 1], a, bb[2], bb[3]);
          }
        }
      }
    }
  }

  // ---- finalize
  if (!warp_live) return;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = l_run[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    l_run[r] = l;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = row_g[r];
    if (i < p.Tq) {
      const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
      bf16* op = o + (int64_t)b * o_bs + (int64_t)i * o_rs + h * 64;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const uint32_t u = pack_bf16(oacc[dt][2 * r] * inv, oacc[dt][2 * r + 1] * inv);
        *reinterpret_cast<uint32_t*>(op + dt * 8 + 2 * t) = u;
      }
      // natural-log log-sum-exp of the (scaled, biased) scores
      if (t == 0) lse[((int64_t)b * p.H + h) * p.Tq + i] = (l_run[r] > 0.f) ? (m_run[r] + __log2f(l_run[r])) * kLn2 : -INFINITY;
    }
  }
}

// Column sums of one CTA's 64 x 64 gradient tile (rows = its queries or keys, held as 8 x [16 x 8] warp accumulators):
// the bias gradient of the projection that produced q / k / v is the column sum of dq / dk / dv, and the tile is in
// registers here -- one partial row per CTA goes to `out` (64 floats), a small reduction over the CTAs finishes it
// (deterministic; replaces a separate pass that re-read the whole gradient from HBM).  Rows outside the sequence
// hold exact zeros.  Every thread of the CTA must call this.
__device__ __forceinline__ void tile_colsum(const float (*acc)[4], float (*cs)[64], float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 3;
  __syncthreads();  // cs may still be read by a previous call
#pragma unroll
  for (int dt = 0; dt < 8; ++dt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = acc[dt][e] + acc[dt][2 + e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) cs[warp][dt * 8 + 2 * t + e] = v;
    }
  __syncthreads();
  if (threadIdx.x < 64) out[threadIdx.x] = cs[0][threadIdx.x] + cs[1][threadIdx.x] + cs[2][threadIdx.x] + cs[3][threadIdx.x];
}

// ===================================================================================== backward
struct AttnBwdExtra {
  const bf16* d_o;
  int64_t do_bs, do_rs;
  const bf16* o;  // forward output (for delta = rowsum(dO * O))
  int64_t o_bs, o_rs;
  const float* lse;
  float* delta;   // written by the dQ kernel, read by the dK/dV kernel that runs after it
  bf16 *dq, *dk, *dv, *dpq, *dpk;
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  float* dtable;
  float *dq_colsum, *dk_colsum, *dv_colsum;  // [B * tiles, H * 64] per-CTA column sums or NULL
  AttnDrop drop;
};

// ---- dK / dV: one CTA per 64-key tile, loops over query tiles.  Works on transposed scores
// S^T[key, query] so every accumulator row belongs to this CTA's keys.
// Query-side operands (Q, dO, lse, delta) are double-buffered when no table column competes for shared memory
// (NSB = 2: tile qt+1 is fetched while tile qt is consumed, one barrier per tile).
template <bool HAS_POS, bool HAS_TAB, bool DROP = false>
__global__ void __launch_bounds__(128, HAS_POS ? ATTN_BWD_MINB - 1 : ATTN_BWD_MINB) attn_bwd_dkv_kernel(const AttnCommon p, const AttnBwdExtra e) {
  constexpr int NH = HAS_POS ? 2 : 1;
  constexpr int NSB = HAS_TAB ? 1 : 2;
  constexpr int QST = (NH + 1) * TILE_BYTES;      // one query-side stage: Q (NH tiles) | dO
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sK = smem_u32(smem);             // NH tiles (this CTA's keys)
  const uint32_t sV = sK + NH * TILE_BYTES;       // 1 tile
  const uint32_t sQ0 = sV + TILE_BYTES;           // NSB stages
  float* lse_s = reinterpret_cast<float*>(smem + (NH + 1) * TILE_BYTES + NSB * QST);  // [NSB][TILE] lse * log2(e)
  float* dlt_s = lse_s + NSB * TILE;                                                   // [NSB][TILE] delta * scale
  uint32_t* kmask_s = reinterpret_cast<uint32_t*>(dlt_s + NSB * TILE);
  float* tab_s = reinterpret_cast<float*>(kmask_s + 2);

  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int k0 = kb * TILE;
  const bool warp_live = k0 + warp * 16 < p.Tk;
  pdl_launch();
  pdl_wait();  // PDL: no global access above (this kernel reads the delta written by the dQ kernel before it)
  const bf16* qg = p.q + (int64_t)b * p.q_bs + h * 64;
  const bf16* kg = p.k + (int64_t)b * p.k_bs + h * 64;
  const bf16* vg = p.v + (int64_t)b * p.v_bs + h * 64;
  const bf16* pqg = HAS_POS ? p.pq + (int64_t)b * p.pq_bs + h * 64 : nullptr;
  const bf16* pkg = HAS_POS ? p.pk + (int64_t)b * p.pk_bs + h * 64 : nullptr;
  const bf16* dog = e.d_o + (int64_t)b * e.do_bs + h * 64;

  if (HAS_TAB)
    for (int i = threadIdx.x; i < p.n_buckets; i += 128) tab_s[i] = p.table[(int64_t)i * p.H + h] * kLog2e;

  const int n_q = (p.Tq + TILE - 1) / TILE;
  const int q_start = p.causal ? (k0 / TILE) : 0;  // queries i < k0 never see these keys
  auto load_q = [&](int qt, int stg) {
    const int q1 = qt * TILE;
    const uint32_t sQ = sQ0 + stg * QST;
    load_tile_async(sQ, qg, p.q_rs, q1, p.Tq);
    if (HAS_POS) load_tile_async(sQ + TILE_BYTES, pqg, p.pq_rs, q1, p.Tq);
    load_tile_async(sQ + NH * TILE_BYTES, dog, e.do_rs, q1, p.Tq);
    cp_commit();
    if (threadIdx.x < TILE) {
      const int i = q1 + threadIdx.x;
      const float l = i < p.Tq ? e.lse[((int64_t)b * p.H + h) * p.Tq + i] : -INFINITY;
      // rows without any visible key (or past Tq) get a huge finite lse: every probability becomes ex2(-huge) = 0
      lse_s[stg * TILE + threadIdx.x] = (l == -INFINITY) ? 1e30f : l * kLog2e;
      dlt_s[stg * TILE + threadIdx.x] = i < p.Tq ? e.delta[((int64_t)b * p.H + h) * p.Tq + i] * p.scale : 0.f;
    }
  };

  load_tile_async(sK, kg, p.k_rs, k0, p.Tk);
  if (HAS_POS) load_tile_async(sK + TILE_BYTES, pkg, p.pk_rs, k0, p.Tk);
  load_tile_async(sV, vg, p.v_rs, k0, p.Tk);
  build_kmask(p, b, k0, kmask_s);
  if (q_start < n_q) load_q(q_start, 0);  // one commit group together with K / V

  float dk[NH * 8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < NH * 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[i][j] = 0.f;
  const int key_g[2] = {k0 + warp * 16 + g, k0 + warp * 16 + g + 8};
  uint2 dkey = make_uint2(0u, 1u);
  if (DROP) dkey = attn_drop_key(e.drop, b * p.H + h);

  for (int qt = q_start; qt < n_q; ++qt) {
    const int q0 = qt * TILE;
    const int stg = NSB == 2 ? ((qt - q_start) & 1) : 0;
    if (NSB == 2) {
      cp_wait<0>();      // tile qt has landed ...
      __syncthreads();   // ... for every thread; and everyone finished tile qt-1
      if (qt + 1 < n_q) load_q(qt + 1, stg ^ 1);
    } else {
      if (qt > q_start) {
        __syncthreads();  // previous iteration finished reading the query-side stage
        load_q(qt, 0);
      }
      cp_wait<0>();
      __syncthreads();
    }
    if (!warp_live) continue;
    const uint32_t sQ = sQ0 + stg * QST, sDO = sQ + NH * TILE_BYTES;
    const float* lse2 = lse_s + stg * TILE;
    const float* dls = dlt_s + stg * TILE;

    const int nq = min(TILE, p.Tq - q0);
    const int np_n = (nq + 15) >> 4;  // 16-query groups holding a valid query
    const uint64_t kmask = (uint64_t)kmask_s[0] | ((uint64_t)kmask_s[1] << 32);

    uint32_t ip[8][2];
    if (HAS_TAB) load_idx_pairs<true>(p, q0, k0, warp * 16, ip);
    // S^T[key, query] = K Q^T (+ PK PQ^T)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      const uint32_t tk = sK + hf * TILE_BYTES, tq = sQ + hf * TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        frag_a(tk, warp * 16, kk, a);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          if (np < np_n) {
            uint32_t bb[4];
            frag_b(tq, np * 16, kk, bb);
            mma16816(s[2 * np], a, bb[0], bb[1]);
            mma16816(s[2 * np + 1], a, bb[2], bb[3]);
          }
        }
      }
    }
    // P^T = ex2(S^T * mult - lse2[query])   (masked scores are -inf -> 0)
    const float mult = finish_tile<true, HAS_TAB>(p, s, q0, k0, warp * 16, kmask, tab_s, ip, [](int, int, int) {});
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int el = 0; el < 4; ++el) s[nt][el] = fast_ex2(fmaf(s[nt][el], mult, -lse2[nt * 8 + 2 * t + (el & 1)]));
    if (DROP) {  // dropped probabilities are kept with a NEGATIVE sign: |s| = P, sign = keep bit
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int el = 0; el < 4; ++el) {
          const uint32_t ij = (uint32_t)(q0 + nt * 8 + 2 * t + (el & 1)) * (uint32_t)p.Tk + (uint32_t)key_g[el >> 1];
          if (!attn_keep(dkey, ij, e.drop.thresh32)) s[nt][el] = -s[nt][el];
        }
    }
    const float ik = DROP ? e.drop.inv_keep : 1.0f;
    // dV += P_drop^T dO
#pragma unroll
    for (int kb2 = 0; kb2 < 4; ++kb2) {
      if (kb2 < np_n) {
        uint32_t pa[4];
        if (DROP) {
          pa[0] = pack_bf16(fmaxf(s[2 * kb2][0], 0.f) * ik, fmaxf(s[2 * kb2][1], 0.f) * ik);
          pa[1] = pack_bf16(fmaxf(s[2 * kb2][2], 0.f) * ik, fmaxf(s[2 * kb2][3], 0.f) * ik);
          pa[2] = pack_bf16(fmaxf(s[2 * kb2 + 1][0], 0.f) * ik, fmaxf(s[2 * kb2 + 1][1], 0.f) * ik);
          pa[3] = pack_bf16(fmaxf(s[2 * kb2 + 1][2], 0.f) * ik, fmaxf(s[2 * kb2 + 1][3], 0.f) * ik);
        } else {
          pa[0] = pack_bf16(s[2 * kb2][0], s[2 * kb2][1]);
          pa[1] = pack_bf16(s[2 * kb2][2], s[2 * kb2][3]);
          pa[2] = pack_bf16(s[2 * kb2 + 1][0], s[2 * kb2 + 1][1]);
          pa[3] = pack_bf16(s[2 * kb2 + 1][2], s[2 * kb2 + 1][3]);
        }
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t bb[4];
          frag_bt(sDO, kb2 * 16, dp, bb);
          mma16816(dv[2 * dp], pa, bb[0], bb[1]);
          mma16816(dv[2 * dp + 1], pa, bb[2], bb[3]);
        }
      }
    }
    // dP^T[key, query] = V dO^T
    float dp_[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp_[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      frag_a(sV, warp * 16, kk, a);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < np_n) {
          uint32_t bb[4];
          frag_b(sDO, np * 16, kk, bb);
          mma16816(dp_[2 * np], a, bb[0], bb[1]);
          mma16816(dp_[2 * np + 1], a, bb[2], bb[3]);
        }
      }
    }
    // scale * dS^T = P^T * (scale * dP^T - scale * delta[query])
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int el = 0; el < 4; ++el) {
        if (DROP) {  // dP = keep / (1 - p) * dP_drop
          const float dpm = s[nt][el] > 0.f ? dp_[nt][el] * ik : 0.f;
          s[nt][el] = fabsf(s[nt][el]) * fmaf(dpm, p.scale, -dls[nt * 8 + 2 * t + (el & 1)]);
        } else {
          s[nt][el] *= fmaf(dp_[nt][el], p.scale, -dls[nt * 8 + 2 * t + (el & 1)]);
        }
      }
    // dK' += dS^T Q'
#pragma unroll
    for (int kb2 = 0; kb2 < 4; ++kb2) {
      if (kb2 < np_n) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kb2][0], s[2 * kb2][1]);
        a[1] = pack_bf16(s[2 * kb2][2], s[2 * kb2][3]);
        a[2] = pack_bf16(s[2 * kb2 + 1][0], s[2 * kb2 + 1][1]);
        a[3] = pack_bf16(s[2 * kb2 + 1][2], s[2 * kb2 + 1][3]);
#pragma unroll
        for (int hf = 0; hf < NH; ++hf)
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t bb[4];
            frag_bt(sQ + hf * TILE_BYTES, kb2 * 16, dp, bb);
            mma16816(dk[hf * 8 + 2 * dp], a, bb[0], bb[1]);
            mma16816(dk[hf * 8 + 2 * dp + 1], a, bb[2], bb[3]);
          }
      }
    }
  }
  // ---- store dK (+dPK), dV
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int j = key_g[r];
    if (j < p.Tk) {
      bf16* dkp = e.dk + (int64_t)b * e.dk_bs + (int64_t)j * e.dk_rs + h * 64;
      bf16* dvp = e.dv + (int64_t)b * e.dv_bs + (int64_t)j * e.dv_rs + h * 64;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        *reinterpret_cast<uint32_t*>(dkp + dt * 8 + 2 * t) = pack_bf16(dk[dt][2 * r], dk[dt][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvp + dt * 8 + 2 * t) = pack_bf16(dv[dt][2 * r], dv[dt][2 * r + 1]);
      }
      if (HAS_POS) {
        bf16* dpp = e.dpk + ((int64_t)b * p.Tk + j) * (p.H * 64) + h * 64;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt)
          *reinterpret_cast<uint32_t*>(dpp + dt * 8 + 2 * t) = pack_bf16(dk[(NH - 1) * 8 + dt][2 * r], dk[(NH - 1) * 8 + dt][2 * r + 1]);
      }
    }
  }
  if (e.dk_colsum != nullptr) {
    __shared__ float cs[4][64];
    const int64_t prow = ((int64_t)b * gridDim.x + kb) * (p.H * 64) + h * 64;
    tile_colsum(dk, cs, e.dk_colsum + prow);
    tile_colsum(dv, cs, e.dv_colsum + prow);
  }
}

// ---- dQ (+ dPQ, d table): one CTA per 64-query tile, loops over key tiles (K / V double-buffered when no
// table column competes for shared memory).
template <bool HAS_POS, bool HAS_TAB, bool DROP = false>
__global__ void __launch_bounds__(128, HAS_POS ? ATTN_BWD_MINB - 1 : ATTN_BWD_MINB) attn_bwd_dq_kernel(const AttnCommon p, const AttnBwdExtra e) {
  constexpr int NH = HAS_POS ? 2 : 1;
  constexpr int NSB = HAS_TAB ? 1 : 2;
  constexpr int KST = (NH + 1) * TILE_BYTES;      // one key-side stage: K (NH tiles) | V
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);             // NH
  const uint32_t sDO = sQ + NH * TILE_BYTES;      // 1
  const uint32_t sK0 = sDO + TILE_BYTES;          // NSB stages
  uint32_t* kmask_s = reinterpret_cast<uint32_t*>(smem + (NH + 1) * TILE_BYTES + NSB * KST);
  float* tab_s = reinterpret_cast<float*>(kmask_s + ((p.Tk + 63) >> 6) * 2);
  float* dtab_s = tab_s + p.n_buckets;

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = qb * TILE;
  const bool warp_live = q0 + warp * 16 < p.Tq;
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const bf16* qg = p.q + (int64_t)b * p.q_bs + h * 64;
  const bf16* kg = p.k + (int64_t)b * p.k_bs + h * 64;
  const bf16* vg = p.v + (int64_t)b * p.v_bs + h * 64;
  const bf16* pqg = HAS_POS ? p.pq + (int64_t)b * p.pq_bs + h * 64 : nullptr;
  const bf16* pkg = HAS_POS ? p.pk + (int64_t)b * p.pk_bs + h * 64 : nullptr;
  const bf16* dog = e.d_o + (int64_t)b * e.do_bs + h * 64;
  constexpr bool has_tab = HAS_TAB;

  auto load_kv = [&](int kvi, int stg) {
    const int k1 = kvi * TILE;
    const uint32_t sK = sK0 + stg * KST;
    load_tile_async(sK, kg, p.k_rs, k1, p.Tk);
    if (HAS_POS) load_tile_async(sK + TILE_BYTES, pkg, p.pk_rs, k1, p.Tk);
    load_tile_async(sK + NH * TILE_BYTES, vg, p.v_rs, k1, p.Tk);
    cp_commit();
  };
  load_tile_async(sQ, qg, p.q_rs, q0, p.Tq);
  if (HAS_POS) load_tile_async(sQ + TILE_BYTES, pqg, p.pq_rs, q0, p.Tq);
  load_tile_async(sDO, dog, e.do_rs, q0, p.Tq);
  load_kv(0, 0);  // one commit group together with Q / dO

  if (has_tab)
    for (int i = threadIdx.x; i < p.n_buckets; i += 128) {
      tab_s[i] = p.table[(int64_t)i * p.H + h] * kLog2e;
      dtab_s[i] = 0.f;
    }
  build_kmask_all(p, b, kmask_s);

  const int row_g[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};
  float lse2_r[2], dl_r[2], dls_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    // delta[i] = sum_d dO[i,d] * O[i,d]: the 4 lanes of a quad split the 64 columns of row i (16 each)
    const int i = row_g[r];
    float dsum = 0.f;
    if (i < p.Tq) {
      const bf16* dop = dog + (int64_t)i * e.do_rs + t * 16;
      const bf16* op = e.o + (int64_t)b * e.o_bs + (int64_t)i * e.o_rs + h * 64 + t * 16;
#pragma unroll
      for (int c = 0; c < 16; c += 8) {
        const f8 x = load8(dop + c), y = load8(op + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) dsum += x.v[j] * y.v[j];
      }
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
    const float l = i < p.Tq ? e.lse[((int64_t)b * p.H + h) * p.Tq + i] : -INFINITY;
    lse2_r[r] = (l == -INFINITY) ? 1e30f : l * kLog2e;  // no visible key / past Tq: every probability becomes 0
    dl_r[r] = dsum;
    dls_r[r] = dsum * p.scale;
    if (i < p.Tq && t == 0) e.delta[((int64_t)b * p.H + h) * p.Tq + i] = dsum;  // for the dK/dV kernel
  }
  float dq[NH * 8][4];
#pragma unroll
  for (int i = 0; i < NH * 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  uint2 dkey = make_uint2(0u, 1u);
  if (DROP) dkey = attn_drop_key(e.drop, b * p.H + h);

  const int n_kv = p.causal ? min((p.Tk + TILE - 1) / TILE, (q0 + TILE + TILE - 1) / TILE) : (p.Tk + TILE - 1) / TILE;
  for (int kv = 0; kv < n_kv; ++kv) {
    const int k0 = kv * TILE;
    const int stg = NSB == 2 ? (kv & 1) : 0;
    if (NSB == 2) {
      cp_wait<0>();      // tile kv has landed ...
      __syncthreads();   // ... for every thread; and everyone finished tile kv-1 (first time: bitmap / table staged)
      if (kv + 1 < n_kv) load_kv(kv + 1, stg ^ 1);
    } else {
      if (kv > 0) {
        __syncthreads();
        load_kv(kv, 0);
      }
      cp_wait<0>();
      __syncthreads();
    }
    if (!warp_live) continue;
    const uint32_t sK = sK0 + stg * KST, sV = sK + NH * TILE_BYTES;

    const int nk = min(TILE, p.Tk - k0);
    const int np_n = (nk + 15) >> 4;
    const uint64_t kmask = (uint64_t)kmask_s[2 * kv] | ((uint64_t)kmask_s[2 * kv + 1] << 32);
    uint32_t ip[8][2];
    if (HAS_TAB) load_idx_pairs<false>(p, q0, k0, warp * 16, ip);

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      const uint32_t tq = sQ + hf * TILE_BYTES, tk = sK + hf * TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        frag_a(tq, warp * 16, kk, a);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          if (np < np_n) {
            uint32_t bb[4];
            frag_b(tk, np * 16, kk, bb);
            mma16816(s[2 * np], a, bb[0], bb[1]);
            mma16816(s[2 * np + 1], a, bb[2], bb[3]);
          }
        }
      }
    }
    // dP = dO V^T
    float dp_[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dp_[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      frag_a(sDO, warp * 16, kk, a);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        if (np < np_n) {
          uint32_t bb[4];
          frag_b(sV, np * 16, kk, bb);
          mma16816(dp_[2 * np], a, bb[0], bb[1]);
          mma16816(dp_[2 * np + 1], a, bb[2], bb[3]);
        }
      }
    }
    // P, dS ; relative-position table gradient
    int idxs[has_tab ? 8 : 1][4];
    if (has_tab) {
#pragma unroll
      for (int nt = 0; nt < (has_tab ? 8 : 1); ++nt)
#pragma unroll
        for (int el = 0; el < 4; ++el) idxs[nt][el] = -1;
    }
    const float mult = finish_tile<false, HAS_TAB>(p, s, q0, k0, warp * 16, kmask, tab_s, ip, [&](int nt, int el, int idx) { if (has_tab) idxs[has_tab ? nt : 0][el] = idx; });
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int el = 0; el < 4; ++el) {
        const float pe = fast_ex2(fmaf(s[nt][el], mult, -lse2_r[el >> 1]));  // masked scores are -inf -> 0
        if (DROP) {  // dP = keep / (1 - p) * dP_drop
          const uint32_t ij = (uint32_t)row_g[el >> 1] * (uint32_t)p.Tk + (uint32_t)(k0 + nt * 8 + 2 * t + (el & 1));
          dp_[nt][el] = attn_keep(dkey, ij, e.drop.thresh32) ? dp_[nt][el] * e.drop.inv_keep : 0.f;
        }
        if (has_tab) {
          const float ds = pe * (dp_[nt][el] - dl_r[el >> 1]);
          const int ix = idxs[has_tab ? nt : 0][el];
          if (ix >= 0 && ds != 0.f) atomicAdd(dtab_s + ix, ds);
          s[nt][el] = ds * p.scale;
        } else {
          s[nt][el] = pe * fmaf(dp_[nt][el], p.scale, -dls_r[el >> 1]);  // scale * dS
        }
      }
    // dQ' += dS K'
#pragma unroll
    for (int kb2 = 0; kb2 < 4; ++kb2) {
      if (kb2 < np_n) {
        uint32_t a[4];
        a[0] = pack_bf16(s[2 * kb2][0], s[2 * kb2][1]);
        a[1] = pack_bf16(s[2 * kb2][2], s[2 * kb2][3]);
        a[2] = pack_bf16(s[2 * kb2 + 1][0], s[2 * kb2 + 1][1]);
        a[3] = pack_bf16(s[2 * kb2 + 1][2], s[2 * kb2 + 1][3]);
#pragma unroll
        for (int hf = 0; hf < NH; ++hf)
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            uint32_t bb[4];
            frag_bt(sK + hf * TILE_BYTES, kb2 * 16, dp, bb);
            mma16816(dq[hf * 8 + 2 * dp], a, bb[0], bb[1]);
            mma16816(dq[hf * 8 + 2 * dp + 1], a, bb[2], bb[3]);
          }
      }
    }
  }
  if (warp_live) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = row_g[r];
      if (i < p.Tq) {
        bf16* dqp = e.dq + (int64_t)b * e.dq_bs + (int64_t)i * e.dq_rs + h * 64;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<uint32_t*>(dqp + dt * 8 + 2 * t) = pack_bf16(dq[dt][2 * r], dq[dt][2 * r + 1]);
        if (HAS_POS) {
          bf16* dpp = e.dpq + ((int64_t)b * p.Tq + i) * (p.H * 64) + h * 64;
#pragma unroll
          for (int dt = 0; dt < 8; ++dt)
            *reinterpret_cast<uint32_t*>(dpp + dt * 8 + 2 * t) = pack_bf16(dq[(NH - 1) * 8 + dt][2 * r], dq[(NH - 1) * 8 + dt][2 * r + 1]);
        }
      }
    }
  }
  if (has_tab && e.dtable != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.n_buckets; i += 128) {
      const float v = dtab_s[i];
      if (v != 0.f) atomicAdd(e.dtable + (int64_t)i * p.H + h, v);
    }
  }
  if (e.dq_colsum != nullptr) {
    __shared__ float cs[4][64];
    tile_colsum(dq, cs, e.dq_colsum + ((int64_t)b * gridDim.x + qb) * (p.H * 64) + h * 64);
  }
}

int fill_common(const ofab_attn_fwd_args* a, AttnCommon& c) {
  OFAB_REQUIRE(a->B > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, "ofab_attn: empty problem");
  OFAB_REQUIRE(a->q && a->k && a->v, "ofab_attn: q/k/v NULL");
  OFAB_REQUIRE((a->pq == nullptr) == (a->pk == nullptr), "ofab_attn: pq and pk must both be given or both NULL");
  OFAB_REQUIRE((a->rp_idx == nullptr) == (a->table == nullptr), "ofab_attn: rp_idx and table must both be given or both NULL");
  OFAB_REQUIRE(a->rp_idx == nullptr || (a->n_buckets > 0 && a->n_buckets <= 22900), "ofab_attn: n_buckets=%d out of range (1..22900: one fp32 table column (+ its gradient) per head must fit in shared memory)", a->n_buckets);
  OFAB_REQUIRE(a->q_rs % 8 == 0 && a->k_rs % 8 == 0 && a->v_rs % 8 == 0 && a->q_bs % 8 == 0 && a->k_bs % 8 == 0 && a->v_bs % 8 == 0,
               "ofab_attn: q/k/v strides must be multiples of 8 elements (16-byte rows)");
  OFAB_REQUIRE(a->pq == nullptr || (a->pq_rs % 8 == 0 && a->pk_rs % 8 == 0 && a->pq_bs % 8 == 0 && a->pk_bs % 8 == 0),
               "ofab_attn: pq/pk strides must be multiples of 8 elements");
  c.B = a->B; c.H = a->H; c.Tq = a->Tq; c.Tk = a->Tk;
  c.q = (const bf16*)a->q; c.k = (const bf16*)a->k; c.v = (const bf16*)a->v;
  c.pq = (const bf16*)a->pq; c.pk = (const bf16*)a->pk;
  c.q_bs = a->q_bs; c.q_rs = a->q_rs; c.k_bs = a->k_bs; c.k_rs = a->k_rs; c.v_bs = a->v_bs; c.v_rs = a->v_rs;
  c.pq_bs = a->pq_bs; c.pq_rs = a->pq_rs; c.pk_bs = a->pk_bs; c.pk_rs = a->pk_rs;
  OFAB_REQUIRE(a->rp_idx == nullptr || (a->rp_ld >= a->Tk && a->rp_ld % 2 == 0 && (((uintptr_t)a->rp_idx) & 3) == 0),
               "ofab_attn: rp_idx needs an even row length rp_ld=%d >= Tk and 4-byte alignment", a->rp_ld);
  OFAB_REQUIRE(a->rp_idx_t == nullptr || (a->rp_ld_t >= a->Tq && a->rp_ld_t % 2 == 0 && (((uintptr_t)a->rp_idx_t) & 3) == 0),
               "ofab_attn: rp_idx_t needs an even row length rp_ld_t=%d >= Tq and 4-byte alignment", a->rp_ld_t);
  c.rp_idx = a->rp_idx; c.rp_idx_t = a->rp_idx_t; c.rp_ld = a->rp_ld; c.rp_ld_t = a->rp_ld_t;
  c.table = a->table; c.n_buckets = a->rp_idx ? a->n_buckets : 0;
  c.kpm = a->kpm; c.causal = a->causal; c.scale = a->scale;
  return OFAB_OK;
}

// optional attention dropout -> kernel form; `on` = descriptor given and p > 0
int fill_drop(const ofab_dropout* d, AttnDrop& ad, bool& on, int Tq, int Tk) {
  ad = AttnDrop{nullptr, 0u, 0u, 1.0f};
  on = d != nullptr && d->p > 0.f;
  if (!on) return OFAB_OK;
  OFAB_REQUIRE(d->state != nullptr && d->p < 1.f && d->drop_path == 0.f, "ofab_attn: bad attention dropout descriptor (p=%g drop_path=%g)", (double)d->p, (double)d->drop_path);
  OFAB_REQUIRE((int64_t)Tq * Tk < (1ll << 32), "ofab_attn: Tq * Tk must fit 32 bits with dropout");
  ad.state = reinterpret_cast<const unsigned long long*>(d->state);
  ad.site = d->site;
  const double t = (double)d->p * 4294967296.0;
  ad.thresh32 = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
  ad.inv_keep = 1.0f / (1.0f - d->p);
  return OFAB_OK;
}

template <typename K>
int set_smem(K kern, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return ofab_cuda_fail(e, what);
  return OFAB_OK;
}

}  // namespace

extern "C" int ofab_attn_fwd(const ofab_attn_fwd_args* a, ofab_stream_t stream) {
  AttnCommon c;
  int rc = fill_common(a, c);
  if (rc) return rc;
  OFAB_REQUIRE(a->o && a->lse, "ofab_attn_fwd: o/lse NULL");
  OFAB_REQUIRE(a->o_rs % 8 == 0 && a->o_bs % 8 == 0 && ((uintptr_t)a->o & 15) == 0, "ofab_attn_fwd: o must have 16-byte aligned rows");
  if (ofab_attn_tc_eligible(a)) {  // tcgen05 kernels (attn_tc.cu): everything without structured position terms
    rc = ofab_attn_tc_fwd(a, stream);
    if (rc <= 0) return rc;
  }
  OFAB_REQUIRE(a->bias == nullptr, "ofab_attn_fwd: a dense bias tile needs the tcgen05 kernels (no pq / pk / rp_idx, 16-byte aligned strides)");
  const bool pos = a->pq != nullptr, tab = a->rp_idx != nullptr;
  const int nh = pos ? 2 : 1;
  const int kwords = ((a->Tk + 63) / 64) * 2;  // key-validity bitmap
  const int smem = (3 * nh + 2) * TILE_BYTES + kwords * 4 + c.n_buckets * 4;
  dim3 grid((a->Tq + TILE - 1) / TILE, a->H, a->B);
  cudaStream_t st = (cudaStream_t)stream;
  AttnDrop ad;
  bool drop_on;
  if ((rc = fill_drop(a->drop, ad, drop_on, a->Tq, a->Tk))) return rc;
#define FWD_(P, T, D)                                                                                     \
  {                                                                                                       \
    if ((rc = set_smem(attn_fwd_kernel<P, T, D>, smem, "ofab_attn_fwd smem"))) return rc;                 \
    ofab_launch(attn_fwd_kernel<P, T, D>, grid, dim3(128), smem, st, c, (bf16*)a->o, a->o_bs, a->o_rs, a->lse, ad); \
  }
#define FWD(P, T) { if (drop_on) FWD_(P, T, true) else FWD_(P, T, false) }
  if (pos && tab) FWD(true, true) else if (pos) FWD(true, false) else if (tab) FWD(false, true) else FWD(false, false)
#undef FWD
#undef FWD_
  OFAB_LAUNCH_CHECK("ofab_attn_fwd");
  return OFAB_OK;
}

extern "C" int ofab_attn_bwd(const ofab_attn_bwd_args* a, ofab_stream_t stream) {
  AttnCommon c;
  int rc = fill_common(&a->f, c);
  if (rc) return rc;
  OFAB_REQUIRE(a->d_o && a->dq && a->dk && a->dv && a->delta && a->f.o && a->f.lse, "ofab_attn_bwd: NULL tensor");
  const bool pos = a->f.pq != nullptr, tab = a->f.rp_idx != nullptr;
  OFAB_REQUIRE(!pos || (a->dpq && a->dpk), "ofab_attn_bwd: dpq/dpk required when pq/pk are given");
  OFAB_REQUIRE(!tab || a->f.rp_idx_t != nullptr, "ofab_attn_bwd: rp_idx_t (the transposed bucket map) is required with rp_idx");
  OFAB_REQUIRE(a->dq_rs % 2 == 0 && a->dk_rs % 2 == 0 && a->dv_rs % 2 == 0, "ofab_attn_bwd: grad strides must be even");
  OFAB_REQUIRE(a->do_rs % 8 == 0 && a->do_bs % 8 == 0 && a->f.o_rs % 8 == 0 && a->f.o_bs % 8 == 0, "ofab_attn_bwd: dO / O strides must be multiples of 8");
  if (ofab_attn_tc_eligible(&a->f) && a->dq_rs % 8 == 0 && a->dk_rs % 8 == 0 && a->dv_rs % 8 == 0 ) {
    rc = ofab_attn_tc_bwd(a, stream);
    if (rc <= 0) return rc;
  }
  OFAB_REQUIRE(a->f.bias == nullptr, "ofab_attn_bwd: a dense bias tile needs the tcgen05 kernels");
  cudaStream_t st = (cudaStream_t)stream;
  AttnBwdExtra e;
  e.d_o = (const bf16*)a->d_o; e.do_bs = a->do_bs; e.do_rs = a->do_rs;
  e.o = (const bf16*)a->f.o; e.o_bs = a->f.o_bs; e.o_rs = a->f.o_rs;
  e.lse = a->f.lse; e.delta = a->delta;
  e.dq = (bf16*)a->dq; e.dk = (bf16*)a->dk; e.dv = (bf16*)a->dv; e.dpq = (bf16*)a->dpq; e.dpk = (bf16*)a->dpk;
  e.dq_bs = a->dq_bs; e.dq_rs = a->dq_rs; e.dk_bs = a->dk_bs; e.dk_rs = a->dk_rs; e.dv_bs = a->dv_bs; e.dv_rs = a->dv_rs;
  e.dtable = a->dtable;
  OFAB_REQUIRE((a->dk_colsum == nullptr) == (a->dv_colsum == nullptr), "ofab_attn_bwd: dk_colsum and dv_colsum go together");
  e.dq_colsum = a->dq_colsum; e.dk_colsum = a->dk_colsum; e.dv_colsum = a->dv_colsum;
  bool drop_on;
  if ((rc = fill_drop(a->f.drop, e.drop, drop_on, a->f.Tq, a->f.Tk))) return rc;
  const int nh = pos ? 2 : 1;
  const int nsb = tab ? 1 : 2;  // query- / key-side stages (double-buffered unless a table column needs the space)
  const int kwords = ((a->f.Tk + 63) / 64) * 2;
  const int smem_kv = (nh + 1) * (1 + nsb) * TILE_BYTES + nsb * 2 * TILE * 4 + 8 + c.n_buckets * 4;
  const int smem_q = (nh + 1) * (1 + nsb) * TILE_BYTES + kwords * 4 + 2 * c.n_buckets * 4;
  dim3 gkv((a->f.Tk + TILE - 1) / TILE, a->f.H, a->f.B), gq((a->f.Tq + TILE - 1) / TILE, a->f.H, a->f.B);
  // dQ first: it also computes delta = rowsum(dO * O) that the dK/dV kernel needs
#define BWD_(P, T, D)                                                                                  \
  {                                                                                                    \
    if ((rc = set_smem(attn_bwd_dq_kernel<P, T, D>, smem_q, "ofab_attn_bwd smem"))) return rc;         \
    if ((rc = set_smem(attn_bwd_dkv_kernel<P, T, D>, smem_kv, "ofab_attn_bwd smem"))) return rc;       \
    ofab_launch(attn_bwd_dq_kernel<P, T, D>, gq, dim3(128), smem_q, st, c, e);                         \
    OFAB_LAUNCH_CHECK("ofab_attn_bwd dq");                                                             \
    ofab_launch(attn_bwd_dkv_kernel<P, T, D>, gkv, dim3(128), smem_kv, st, c, e);                      \
  }
#define BWD(P, T) { if (drop_on) BWD_(P, T, true) else BWD_(P, T, false) }
  if (pos && tab) BWD(true, true) else if (pos) BWD(true, false) else if (tab) BWD(false, true) else BWD(false, false)
#undef BWD
#undef BWD_
  OFAB_LAUNCH_CHECK("ofab_attn_bwd dkv");
  return OFAB_OK;
}
