// LayerNorm family (HBM-bound): plain / GELU-fused LN, the fused LN -> +residual -> LN junction,
// and the column reductions that finish dgamma/dbeta and bias gradients.
//
// Layout: one row is owned by cols/8 threads (rounded up to whole warps), one 8-column vector per thread; a CTA of
// ~384 threads works on a GROUP of blockDim/tpr consecutive rows per iteration.  Every kernel is persistent (grid =
// 2 CTAs per SM) and reads its inputs through an ASYNC ROW PIPELINE: the contiguous bytes of the next row groups are
// fetched by the TMA engine (cp.async.bulk, 1-D) into a shared-memory ring while the current group is being
// reduced, so the bytes in flight per SM (3-4 groups x 6-36 KB per CTA) do not depend on register-held loads or on
// occupancy.  Threads take their 16/32 bytes from shared memory, results leave with plain vector stores.
// Backward kernels keep per-column dgamma/dbeta partial sums in registers over their whole row range and emit one
// partial row per CTA (OFAB_LN_PARTIAL_ROWS rows, finished by reduce_partials).
#include <mutex>
#include <unordered_set>

#include "common.cuh"

#define OFAB_LN_PARTIAL_ROWS 296  // 2 x 148 SMs: persistent backward grid = resident CTAs

extern "C" int ofab_ln_partial_rows(void) { return OFAB_LN_PARTIAL_ROWS; }

namespace {

// Thread layout of every LayerNorm kernel: a row is owned by `tpr` threads (a multiple of 32), each holding ONE
// vector of 8 consecutive columns; a block of rpb * tpr threads works on rpb rows at a time.
struct RowCtx {
  int tpr, rpb, rib, lane_in_row, wir, wpr, c;
  bool col_ok;
};
__device__ __forceinline__ RowCtx row_ctx(int tpr, int cols) {
  RowCtx r;
  r.tpr = tpr;
  r.rpb = blockDim.x / tpr;
  r.rib = threadIdx.x / tpr;
  r.lane_in_row = threadIdx.x % tpr;
  r.wpr = tpr >> 5;
  r.wir = r.lane_in_row >> 5;
  r.c = r.lane_in_row * 8;
  r.col_ok = r.c < cols && r.rib < r.rpb;
  return r;
}

// Sums over the threads of one row group; all threads of the block call these together.  Cross-warp step: one
// partial per warp goes to shared memory (slot [row group][warp in row], 16 slots per group, unused slots stay zero
// from red_init), then every thread adds the partials of its row, read directly with 16-byte loads in a fixed order (so
// every thread of the row gets the same bits).  This replaced a second 16-lane shuffle butterfly: the 768 / 1024 column
// kernels are bound by the latency of exactly this chain (ncu: short-scoreboard stalls), and keeping both forms
// in one kernel cost registers.  `red` is double-buffered ([2][256] values) so ONE __syncthreads per call suffices: a
// buffer is rewritten two calls later, after every thread has passed the barrier in between.
// `sync_always`: barrier even when rows are one warp wide (callers use it to release a pipeline stage).
template <typename T>
__device__ __forceinline__ void red_init(T* red) {  // call before the first __syncthreads of the kernel
  for (int i = threadIdx.x; i < 512; i += blockDim.x) red[i] = T{};
}
// NARROW (rows at most 4 warps wide, chosen at dispatch): two 16-byte loads of the row's partials; otherwise the 16-lane
// butterfly.  A compile-time choice: both forms in one kernel cost registers (spills) in kernels that sit at their cap.
template <bool NARROW>
__device__ __forceinline__ void group_sum2(float& a, float& b, float2* red, int& flip, const RowCtx& r, bool sync_always) {
  a = warp_sum(a);
  b = warp_sum(b);
  if (r.wpr == 1) {
    if (sync_always) __syncthreads();
    return;
  }
  float2* buf = red + flip * 256;
  flip ^= 1;
  if ((threadIdx.x & 31) == 0) buf[r.rib * 16 + r.wir] = make_float2(a, b);
  __syncthreads();
  if (NARROW) {
    const float4 p = *reinterpret_cast<const float4*>(buf + r.rib * 16), q = *reinterpret_cast<const float4*>(buf + r.rib * 16 + 2);
    a = (p.x + p.z) + (q.x + q.z);
    b = (p.y + p.w) + (q.y + q.w);
  } else {
    float2 s2 = buf[r.rib * 16 + (threadIdx.x & 15)];  // slots past the row's warp count are zero (red_init)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s2.x += __shfl_xor_sync(0xffffffffu, s2.x, o);
      s2.y += __shfl_xor_sync(0xffffffffu, s2.y, o);
    }
    a = s2.x;
    b = s2.y;
  }
}

// ---- persistent schedule + async input ring --------------------------------------------------------------
// Iteration `it` of a CTA works on row group g = blockIdx.x + it * gridDim.x (rows g*rpb .. g*rpb+rpb-1, contiguous
// in memory).  Dynamic shared memory: [128 B of mbarriers][NST stages]; a stage holds, for each of the NIN inputs,
// the rpb rows of the group back to back.
constexpr int kLnBarBytes = 128;
template <int NIN>
struct RowInputs {
  const uint8_t* p[NIN];
  uint32_t row_bytes[NIN];  // cols * sizeof(element)
  uint32_t off[NIN];        // byte offset of input k inside a stage (rpb * sum of earlier row_bytes)
  uint32_t stage_bytes;
};
template <int NIN>
__device__ __forceinline__ void ring_issue(const RowInputs<NIN>& in, uint8_t* ring, uint64_t* bars, int stage, int64_t it, int64_t rows, int rpb) {
  const int64_t row0 = ((int64_t)blockIdx.x + it * gridDim.x) * rpb;
  const int64_t left = rows - row0;
  const uint32_t nr = (uint32_t)(left < rpb ? left : rpb);
  uint32_t total = 0;
#pragma unroll
  for (int k = 0; k < NIN; ++k) total += nr * in.row_bytes[k];
  rowpipe::expect(bars + stage, total);
  uint8_t* dst = ring + (size_t)stage * in.stage_bytes;
#pragma unroll
  for (int k = 0; k < NIN; ++k) rowpipe::load(dst + in.off[k], in.p[k] + row0 * in.row_bytes[k], nr * in.row_bytes[k], bars + stage);
}
// common prologue: barrier init + first NST groups in flight.  Returns this CTA's iteration count.
template <int NIN, int NST>
__device__ __forceinline__ int64_t ring_start(const RowInputs<NIN>& in, uint8_t* ring, uint64_t* bars, int64_t rows, int rpb) {
  const int64_t ngroups = (rows + rpb - 1) / rpb;
  const int64_t my_n = (int64_t)blockIdx.x < ngroups ? (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) rowpipe::init(bars, NST);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < NST && i < my_n; ++i) ring_issue<NIN>(in, ring, bars, i, i, rows, rpb);
  return my_n;
}

// Write this block's per-column partial sums (one f8 per thread and slab) as row blockIdx.x of every slab,
// first combining the row groups of the block through shared memory.
template <int NS>
__device__ __forceinline__ void flush_partials(f8 (&acc)[NS], float* __restrict__ partial, int cols, const RowCtx& r, float* buf /* blockDim*8 floats */) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    float* dst = partial + ((int64_t)s * OFAB_LN_PARTIAL_ROWS + blockIdx.x) * cols;
    if (r.rpb == 1) {
      if (r.col_ok) store8(dst + r.c, acc[s]);
    } else {
      __syncthreads();
      if (r.rib < r.rpb) store8(buf + (r.rib * r.tpr + r.lane_in_row) * 8, acc[s]);
      __syncthreads();
      for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float t = 0.f;
        for (int g = 0; g < r.rpb; ++g) t += buf[g * r.tpr * 8 + c];
        dst[c] = t;
      }
    }
  }
}

#define LN_BOUNDS(MAXT) __launch_bounds__(MAXT, (MAXT) > 384 ? 1 : 2)  // MAXT = 288 / 384: two CTAs per SM; 512: one

// ------------------------------------------------------------------------------------ forward
// DROP (GELU form only): activation dropout between GELU and the LayerNorm (transformer_layer.py:195).
// VPT: 8-column vectors per thread.  2 for the wide GELU rows (3072 / 4096 columns: `tpr` = cols / 16 threads per row, the
// thread's second vector sits tpr * 8 columns to the right): the per-iteration overhead (reduction, pipeline wait, loop) is
// paid once per 16 elements, which matters because this kernel is issue-bound (ncu: profiles/r02_ncu_ln_before_packed.csv).
template <typename TX, typename TY, bool GELU, int MAXT, bool DROP = false, int VPT = 1, bool NARROW = false>
__global__ void LN_BOUNDS(MAXT) ln_fwd_kernel(const TX* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                              TY* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int cols,
                                              float eps, int tpr, const DropArgs da) {
  constexpr int NST = 6;
  extern __shared__ __align__(128) uint8_t dsm[];
  pdl_launch();  // the next kernel's CTAs may become resident once all of ours have started ...
  pdl_wait();    // ... and we touch global memory only after the previous kernel has completed
  __shared__ __align__(16) float2 red[512];
  red_init(red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm);
  uint8_t* ring = dsm + kLnBarBytes;
  const RowCtx r = row_ctx(tpr, cols);
  RowInputs<1> in;
  in.p[0] = reinterpret_cast<const uint8_t*>(x);
  in.row_bytes[0] = (uint32_t)cols * sizeof(TX);
  in.off[0] = 0;
  in.stage_bytes = (uint32_t)r.rpb * in.row_bytes[0];
  const int64_t my_n = ring_start<1, NST>(in, ring, bars, rows, r.rpb);
  const float inv_n = 1.0f / (float)cols;
  DropCtx dk;
  if (DROP) dk = drop_ctx(da);
  const int vstep = tpr * 8;  // columns between a thread's vectors
  bool ok[VPT];
  f8 g[VPT], b[VPT];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    ok[k] = r.rib < r.rpb && r.c + k * vstep < cols;
    if (ok[k]) {
      g[k] = load8(gamma + r.c + k * vstep);
      b[k] = load8(beta + r.c + k * vstep);
    }
  }
  int stage = 0, flip = 0;
  uint32_t phase = 0;
  // row and element offset of this thread advance by a constant per iteration (no 64-bit multiplies in the loop)
  const int64_t row_step = (int64_t)gridDim.x * r.rpb, eoff_step = row_step * cols;
  int64_t row = (int64_t)blockIdx.x * r.rpb + r.rib - row_step, eoff = row * cols + r.c;
  for (int64_t it = 0; it < my_n; ++it) {
    row += row_step;
    eoff += eoff_step;
    const bool row_live = row < rows;
    rowpipe::wait(bars + stage, phase);
    f8 v[VPT];
    float s = 0.f, q = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      if (ok[k] && row_live) {
        v[k] = load8(reinterpret_cast<const TX*>(ring + (size_t)stage * in.stage_bytes) + r.rib * cols + r.c + k * vstep);
        if (GELU) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) gelu2(v[k].v[j], v[k].v[j + 1]);
        }
        if (DROP) {
          const f8 m = drop_mask8(da, dk, row, r.c + k * vstep);
#pragma unroll
          for (int j = 0; j < 8; j += 2) mul2(v[k].v[j], v[k].v[j + 1], v[k].v[j], v[k].v[j + 1], m.v[j], m.v[j + 1]);
        }
        // one pass: sum and sum of squares (fp32; the variance is E[v^2] - mean^2, clamped at 0)
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          add2(s, s1, s, s1, v[k].v[j], v[k].v[j + 1]);
          fma2(q, q1, v[k].v[j], v[k].v[j + 1], v[k].v[j], v[k].v[j + 1], q, q1);
        }
      }
    }
    s += s1;
    q += q1;
    group_sum2<NARROW>(s, q, red, flip, r, true);  // barrier: every thread has consumed the stage
    if (threadIdx.x == 0 && it + NST < my_n) ring_issue<1>(in, ring, bars, stage, it + NST, rows, r.rpb);
    const float mu = s * inv_n;
    const float rs = rsqrtf(fmaxf(fmaf(-mu, mu, q * inv_n), 0.f) + eps);
    if (ok[0] && row_live && r.lane_in_row == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      if (ok[k] && row_live) {
        f8 o;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float t0, t1;
          add2(t0, t1, v[k].v[j], v[k].v[j + 1], -mu, -mu);
          mul2(t0, t1, t0, t1, rs, rs);
          fma2(o.v[j], o.v[j + 1], t0, t1, g[k].v[j], g[k].v[j + 1], b[k].v[j], b[k].v[j + 1]);
        }
        store8(y + eoff + k * vstep, o);
      }
    }
    if (++stage == NST) {
      stage = 0;
      phase ^= 1;
    }
  }
}

// ----------------------------------------------------------------------------------- backward
template <typename TDY, typename TX, typename TDX, bool GELU, bool ACCUM, int MAXT, bool DROP = false, bool NARROW = false>
__global__ void LN_BOUNDS(MAXT) ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const bf16* __restrict__ gamma,
                                              const float* __restrict__ mean, const float* __restrict__ rstd, TDX* __restrict__ dx,
                                              float* __restrict__ partial, int64_t rows, int cols, int tpr, const DropArgs da) {
  constexpr int NST = 4;
  extern __shared__ __align__(128) uint8_t dsm[];
  pdl_launch();  // the next kernel's CTAs may become resident once all of ours have started ...
  pdl_wait();    // ... and we touch global memory only after the previous kernel has completed
  __shared__ __align__(16) float2 red[512];
  red_init(red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm);
  uint8_t* ring = dsm + kLnBarBytes;
  const RowCtx r = row_ctx(tpr, cols);
  RowInputs<2> in;
  in.p[0] = reinterpret_cast<const uint8_t*>(x);
  in.p[1] = reinterpret_cast<const uint8_t*>(dy);
  in.row_bytes[0] = (uint32_t)cols * sizeof(TX);
  in.row_bytes[1] = (uint32_t)cols * sizeof(TDY);
  in.off[0] = 0;
  in.off[1] = (uint32_t)r.rpb * in.row_bytes[0];
  in.stage_bytes = (uint32_t)r.rpb * (in.row_bytes[0] + in.row_bytes[1]);
  const int64_t my_n = ring_start<2, NST>(in, ring, bars, rows, r.rpb);
  const float inv_n = 1.0f / (float)cols;
  DropCtx dk;
  if (DROP) dk = drop_ctx(da);
  f8 g;
  if (r.col_ok) g = load8(gamma + r.c);
  f8 acc[3];  // dgamma, dbeta, column sums of dx (= bias gradient of the Linear that produced x)
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s].v[j] = 0.f;
  int stage = 0, flip = 0;
  uint32_t phase = 0;
  // row and element offset of this thread advance by a constant per iteration (no 64-bit multiplies in the loop)
  const int64_t row_step = (int64_t)gridDim.x * r.rpb, eoff_step = row_step * cols;
  int64_t row = (int64_t)blockIdx.x * r.rpb + r.rib - row_step, eoff = row * cols + r.c;
  // the row statistics of the NEXT iteration are requested before this iteration's wait: their L2 latency used to sit at
  // the head of every iteration's dependency chain (ncu source view: 17 % of the stall samples on the first use)
  float nx_mean = 0.f, nx_rstd = 0.f;
  if (r.col_ok && row + row_step < rows) {
    nx_mean = mean[row + row_step];
    nx_rstd = rstd[row + row_step];
  }
  for (int64_t it = 0; it < my_n; ++it) {
    row += row_step;
    eoff += eoff_step;
    const bool live = r.col_ok && row < rows;
    const float cur_mean = nx_mean, cur_rstd = nx_rstd;
    if (r.col_ok && row + row_step < rows) {
      nx_mean = mean[row + row_step];
      nx_rstd = rstd[row + row_step];
    }
    rowpipe::wait(bars + stage, phase);
    float rs = 0.f;
    f8 xh, d, gp;
    float s1 = 0.f, s2 = 0.f;
    if (live) {
      const uint8_t* st = ring + (size_t)stage * in.stage_bytes;
      const f8 pre = load8(reinterpret_cast<const TX*>(st) + r.rib * cols + r.c);
      d = load8(reinterpret_cast<const TDY*>(st + in.off[1]) + r.rib * cols + r.c);
      rs = cur_rstd;
      const float nmr = -cur_mean * rs;
      f8 m;
      if (DROP) m = drop_mask8(da, dk, row, r.c);
      float s1b = 0.f, s2b = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {  // pairs of columns through the packed fp32 pipe
        float a0 = pre.v[j], a1 = pre.v[j + 1];
        if (GELU) {  // one erf evaluation serves both gelu(x) and gelu'(x)
          float c0, c1, p0, p1;
          gelu_parts2(a0, a1, c0, c1, p0, p1);
          add2(gp.v[j], gp.v[j + 1], c0, c1, p0, p1);
          mul2(a0, a1, a0, a1, c0, c1);
          if (DROP) {  // LN input = m * gelu(x): the mask scales the value and the chain-rule factor alike
            mul2(a0, a1, a0, a1, m.v[j], m.v[j + 1]);
            mul2(gp.v[j], gp.v[j + 1], gp.v[j], gp.v[j + 1], m.v[j], m.v[j + 1]);
          }
        }
        fma2(xh.v[j], xh.v[j + 1], a0, a1, rs, rs, nmr, nmr);
        fma2(acc[0].v[j], acc[0].v[j + 1], d.v[j], d.v[j + 1], xh.v[j], xh.v[j + 1], acc[0].v[j], acc[0].v[j + 1]);
        add2(acc[1].v[j], acc[1].v[j + 1], acc[1].v[j], acc[1].v[j + 1], d.v[j], d.v[j + 1]);
        mul2(d.v[j], d.v[j + 1], d.v[j], d.v[j + 1], g.v[j], g.v[j + 1]);
        fma2(s1, s1b, d.v[j], d.v[j + 1], xh.v[j], xh.v[j + 1], s1, s1b);
        add2(s2, s2b, s2, s2b, d.v[j], d.v[j + 1]);
      }
      s1 += s1b;
      s2 += s2b;
    }
    group_sum2<NARROW>(s1, s2, red, flip, r, true);  // barrier: every thread has consumed the stage
    if (threadIdx.x == 0 && it + NST < my_n) ring_issue<2>(in, ring, bars, stage, it + NST, rows, r.rpb);
    if (live) {
      const float c1 = -s1 * inv_n * rs, c2 = -s2 * inv_n * rs;
      f8 o;
      if (ACCUM) o = load8(reinterpret_cast<const TDX*>(dx) + eoff);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float t0, t1;
        fma2(t0, t1, d.v[j], d.v[j + 1], rs, rs, c2, c2);
        fma2(t0, t1, xh.v[j], xh.v[j + 1], c1, c1, t0, t1);  // rs * (d - mean(d) - xh * mean(d xh))
        if (GELU) mul2(t0, t1, t0, t1, gp.v[j], gp.v[j + 1]);
        add2(acc[2].v[j], acc[2].v[j + 1], acc[2].v[j], acc[2].v[j + 1], t0, t1);
        if (ACCUM) {
          add2(o.v[j], o.v[j + 1], o.v[j], o.v[j + 1], t0, t1);
        } else {
          o.v[j] = t0;
          o.v[j + 1] = t1;
        }
      }
      store8(dx + eoff, o);
    }
    if (++stage == NST) {
      stage = 0;
      phase ^= 1;
    }
  }
  flush_partials<3>(acc, partial, cols, r, reinterpret_cast<float*>(ring));
}

// ---------------------------------------------------------------- fused LN -> +res -> LN
// HAS_LN1 = false: x_new = x + a (no first LayerNorm): the deferred residual add of an FFN output fused
// with the next block's pre-LayerNorm.
// DROP: x_new = x + mask * branch  (residual dropout + drop-path of the block, transformer_layer.py:181,87)
template <bool HAS_LN1, int MAXT, bool DROP = false, bool NARROW = false>
__global__ void LN_BOUNDS(MAXT) ln_res_ln_fwd_kernel(const bf16* __restrict__ a, const float* __restrict__ x, const bf16* __restrict__ g1,
                                                     const bf16* __restrict__ b1, const bf16* __restrict__ g2, const bf16* __restrict__ b2,
                                                     float* __restrict__ x_new, bf16* __restrict__ y, float* __restrict__ stats,
                                                     int64_t rows, int cols, float eps, int tpr, const DropArgs da) {
  constexpr int NST = 3;
  extern __shared__ __align__(128) uint8_t dsm[];
  pdl_launch();  // the next kernel's CTAs may become resident once all of ours have started ...
  pdl_wait();    // ... and we touch global memory only after the previous kernel has completed
  __shared__ __align__(16) float2 red[512];
  red_init(red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm);
  uint8_t* ring = dsm + kLnBarBytes;
  const RowCtx r = row_ctx(tpr, cols);
  RowInputs<2> in;
  in.p[0] = reinterpret_cast<const uint8_t*>(x);
  in.p[1] = reinterpret_cast<const uint8_t*>(a);
  in.row_bytes[0] = (uint32_t)cols * 4u;
  in.row_bytes[1] = (uint32_t)cols * 2u;
  in.off[0] = 0;
  in.off[1] = (uint32_t)r.rpb * in.row_bytes[0];
  in.stage_bytes = (uint32_t)r.rpb * (in.row_bytes[0] + in.row_bytes[1]);
  const int64_t my_n = ring_start<2, NST>(in, ring, bars, rows, r.rpb);
  const float inv_n = 1.0f / (float)cols;
  DropCtx dk;
  if (DROP) dk = drop_ctx(da);
  f8 gg1, bb1, gg2, bb2;
  if (r.col_ok) {
    if (HAS_LN1) {
      gg1 = load8(g1 + r.c);
      bb1 = load8(b1 + r.c);
    }
    gg2 = load8(g2 + r.c);
    bb2 = load8(b2 + r.c);
  }
  int stage = 0, flip = 0;
  uint32_t phase = 0;
  // row and element offset of this thread advance by a constant per iteration (no 64-bit multiplies in the loop)
  const int64_t row_step = (int64_t)gridDim.x * r.rpb, eoff_step = row_step * cols;
  int64_t row = (int64_t)blockIdx.x * r.rpb + r.rib - row_step, eoff = row * cols + r.c;
  for (int64_t it = 0; it < my_n; ++it) {
    row += row_step;
    eoff += eoff_step;
    const bool live = r.col_ok && row < rows;
    rowpipe::wait(bars + stage, phase);
    f8 v, xx;
    float s = 0.f, q = 0.f;
    if (live) {
      const uint8_t* st = ring + (size_t)stage * in.stage_bytes;
      xx = load8(reinterpret_cast<const float*>(st) + r.rib * cols + r.c);
      v = load8(reinterpret_cast<const bf16*>(st + in.off[1]) + r.rib * cols + r.c);
      if (HAS_LN1) {  // one pass: sum and sum of squares of the branch
        float s1 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          add2(s, s1, s, s1, v.v[j], v.v[j + 1]);
          fma2(q, q1, v.v[j], v.v[j + 1], v.v[j], v.v[j + 1], q, q1);
        }
        s += s1;
        q += q1;
      }
    }
    // (first) barrier of the iteration: every thread has consumed the stage -> refill it
    float m1 = 0.f, r1 = 1.f;
    if (HAS_LN1) {
      group_sum2<NARROW>(s, q, red, flip, r, true);
      m1 = s * inv_n;
      r1 = rsqrtf(fmaxf(fmaf(-m1, m1, q * inv_n), 0.f) + eps);
    } else {
      __syncthreads();
    }
    if (threadIdx.x == 0 && it + NST < my_n) ring_issue<2>(in, ring, bars, stage, it + NST, rows, r.rpb);
    s = 0.f;
    q = 0.f;
    if (live) {
      f8 m;
      if (DROP) m = drop_mask8(da, dk, row, r.c);
      float s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float b0 = v.v[j], b1v = v.v[j + 1];
        if (HAS_LN1) {
          add2(b0, b1v, b0, b1v, -m1, -m1);
          mul2(b0, b1v, b0, b1v, r1, r1);
          fma2(b0, b1v, b0, b1v, gg1.v[j], gg1.v[j + 1], bb1.v[j], bb1.v[j + 1]);
        }
        if (DROP) mul2(b0, b1v, b0, b1v, m.v[j], m.v[j + 1]);
        add2(v.v[j], v.v[j + 1], xx.v[j], xx.v[j + 1], b0, b1v);
        add2(s, s1, s, s1, v.v[j], v.v[j + 1]);
        fma2(q, q1, v.v[j], v.v[j + 1], v.v[j], v.v[j + 1], q, q1);
      }
      s += s1;
      q += q1;
      store8(x_new + eoff, v);
    }
    group_sum2<NARROW>(s, q, red, flip, r, false);
    const float m2 = s * inv_n;
    const float r2 = rsqrtf(fmaxf(fmaf(-m2, m2, q * inv_n), 0.f) + eps);
    if (live) {
      if (r.lane_in_row == 0) {
        stats[row] = m1;
        stats[rows + row] = r1;
        stats[2 * rows + row] = m2;
        stats[3 * rows + row] = r2;
      }
      f8 o;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float t0, t1;
        add2(t0, t1, v.v[j], v.v[j + 1], -m2, -m2);
        mul2(t0, t1, t0, t1, r2, r2);
        fma2(o.v[j], o.v[j + 1], t0, t1, gg2.v[j], gg2.v[j + 1], bb2.v[j], bb2.v[j + 1]);
      }
      store8(y + eoff, o);
    }
    if (++stage == NST) {
      stage = 0;
      phase ^= 1;
    }
  }
}

template <bool HAS_LN1, int MAXT, bool DROP = false, bool NARROW = false>
__global__ void LN_BOUNDS(MAXT) ln_res_ln_bwd_kernel(const float* __restrict__ dxn, const bf16* __restrict__ dy, const bf16* __restrict__ a,
                                                     const float* __restrict__ x_new, const bf16* __restrict__ g1,
                                                     const bf16* __restrict__ g2, const float* __restrict__ stats,
                                                     float* __restrict__ dxt, bf16* __restrict__ da, float* __restrict__ partial,
                                                     int64_t rows, int cols, int tpr, const DropArgs dra) {
  constexpr int NST = 2;
  constexpr int NIN = HAS_LN1 ? 4 : 3;
  extern __shared__ __align__(128) uint8_t dsm[];
  pdl_launch();  // the next kernel's CTAs may become resident once all of ours have started ...
  pdl_wait();    // ... and we touch global memory only after the previous kernel has completed
  __shared__ __align__(16) float2 red[512];
  red_init(red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm);
  uint8_t* ring = dsm + kLnBarBytes;
  const RowCtx r = row_ctx(tpr, cols);
  RowInputs<NIN> in;
  in.p[0] = reinterpret_cast<const uint8_t*>(x_new);
  in.p[1] = reinterpret_cast<const uint8_t*>(dxn);
  in.p[2] = reinterpret_cast<const uint8_t*>(dy);
  in.row_bytes[0] = in.row_bytes[1] = (uint32_t)cols * 4u;
  in.row_bytes[2] = (uint32_t)cols * 2u;
  in.off[0] = 0;
  in.off[1] = (uint32_t)r.rpb * (uint32_t)cols * 4u;
  in.off[2] = (uint32_t)r.rpb * (uint32_t)cols * 8u;
  if (HAS_LN1) {
    in.p[NIN - 1] = reinterpret_cast<const uint8_t*>(a);
    in.row_bytes[NIN - 1] = (uint32_t)cols * 2u;
    in.off[NIN - 1] = (uint32_t)r.rpb * (uint32_t)cols * 10u;
  }
  in.stage_bytes = (uint32_t)r.rpb * (uint32_t)cols * (HAS_LN1 ? 12u : 10u);
  // Row statistics (mean / rstd of both LayerNorms) travel global -> shared by 4-byte cp.async two iterations ahead of their
  // use, issued by the first thread of each row group.  Loading them at the point of use put an L2 round trip at the head of
  // every iteration's dependency chain (ncu source view: 17 % of the stall samples); prefetching into registers did not fit
  // under the register cap (ptxas parked the values in local memory, whose store waits for the load just the same).
  __shared__ __align__(16) float st_s[3][16][4];
  const bool leader = r.lane_in_row == 0 && r.rib < r.rpb && r.rib < 16;
  auto stats_issue = [&](int slot, int64_t rw) {
    if (rw < rows) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (HAS_LN1 || q >= 2)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rowpipe::s32(&st_s[slot][r.rib][q])), "l"(stats + q * rows + rw) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  {
    const int64_t first = (int64_t)blockIdx.x * r.rpb + r.rib, step = (int64_t)gridDim.x * r.rpb;
    if (leader) {
      stats_issue(0, first);
      stats_issue(1, first + step);
      asm volatile("cp.async.wait_group 1;" ::: "memory");  // row(0)'s are in; ring_start's __syncthreads publishes them
    }
  }
  const int64_t my_n = ring_start<NIN, NST>(in, ring, bars, rows, r.rpb);
  const float inv_n = 1.0f / (float)cols;
  DropCtx dk;
  if (DROP) dk = drop_ctx(dra);
  // gamma vectors stay packed (bf16, 4 registers each) and are unpacked where used: the kernel sits at its register cap and
  // is bound by latency, not by issue slots
  uint4 g1_raw = make_uint4(0u, 0u, 0u, 0u), g2_raw = make_uint4(0u, 0u, 0u, 0u);
  if (r.col_ok) {
    if (HAS_LN1) g1_raw = *reinterpret_cast<const uint4*>(g1 + r.c);
    g2_raw = *reinterpret_cast<const uint4*>(g2 + r.c);
  }
  f8 acc[5];  // dg1, db1, dg2, db2, column sums of da (bias gradient of the Linear that produced a)
#pragma unroll
  for (int s = 0; s < 5; ++s)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s].v[j] = 0.f;
  int stage = 0, flip = 0;
  uint32_t phase = 0;
  // row and element offset of this thread advance by a constant per iteration (no 64-bit multiplies in the loop)
  const int64_t row_step = (int64_t)gridDim.x * r.rpb, eoff_step = row_step * cols;
  int64_t row = (int64_t)blockIdx.x * r.rpb + r.rib - row_step, eoff = row * cols + r.c;
  for (int64_t it = 0; it < my_n; ++it) {
    row += row_step;
    eoff += eoff_step;
    const bool live = r.col_ok && row < rows;
    // the statistics of row(it + 2) start their way into shared memory now; row(it)'s were published by the barrier of the
    // previous iteration (or of the prologue)
    if (leader) stats_issue((int)((it + 2) % 3), row + 2 * row_step);
    const float4 cur = *reinterpret_cast<const float4*>(&st_s[it % 3][r.rib < 16 ? r.rib : 0][0]);
    rowpipe::wait(bars + stage, phase);
    float m1 = 0.f, r1 = 0.f, r2 = 0.f;
    f8 xh, d, tot;
    uint4 a_raw = make_uint4(0u, 0u, 0u, 0u);  // the bf16 row of `a` stays packed until LN1's backward needs it
    float s1 = 0.f, s2 = 0.f;
    if (live) {
      const uint8_t* st = ring + (size_t)stage * in.stage_bytes;
      const int e = r.rib * cols + r.c;
      const f8 xx = load8(reinterpret_cast<const float*>(st) + e);
      tot = load8(reinterpret_cast<const float*>(st + in.off[1]) + e);  // upstream d x_new
      d = load8(reinterpret_cast<const bf16*>(st + in.off[2]) + e);
      if (HAS_LN1) a_raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(st + in.off[NIN - 1]) + e);
      m1 = cur.x;
      r1 = cur.y;
      r2 = cur.w;
      const float nmr = -cur.z * r2;
      float s1b = 0.f, s2b = 0.f;
      const f8 gg2 = unpack8(g2_raw);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        fma2(xh.v[j], xh.v[j + 1], xx.v[j], xx.v[j + 1], r2, r2, nmr, nmr);
        fma2(acc[2].v[j], acc[2].v[j + 1], d.v[j], d.v[j + 1], xh.v[j], xh.v[j + 1], acc[2].v[j], acc[2].v[j + 1]);
        add2(acc[3].v[j], acc[3].v[j + 1], acc[3].v[j], acc[3].v[j + 1], d.v[j], d.v[j + 1]);
        mul2(d.v[j], d.v[j + 1], d.v[j], d.v[j + 1], gg2.v[j], gg2.v[j + 1]);
        fma2(s1, s1b, d.v[j], d.v[j + 1], xh.v[j], xh.v[j + 1], s1, s1b);
        add2(s2, s2b, s2, s2b, d.v[j], d.v[j + 1]);
      }
      s1 += s1b;
      s2 += s2b;
    }
    if (leader) asm volatile("cp.async.wait_group 1;" ::: "memory");  // row(it + 1)'s statistics have landed; the barrier publishes them
    group_sum2<NARROW>(s1, s2, red, flip, r, true);  // barrier: every thread has consumed the stage
    if (threadIdx.x == 0 && it + NST < my_n) ring_issue<NIN>(in, ring, bars, stage, it + NST, rows, r.rpb);
    float t1 = 0.f, t2 = 0.f;
    if (live) {
      const float c1 = -s1 * inv_n * r2, c2 = -s2 * inv_n * r2;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {  // + LN2'(dy)
        float u0, u1;
        fma2(u0, u1, d.v[j], d.v[j + 1], r2, r2, c2, c2);
        fma2(u0, u1, xh.v[j], xh.v[j + 1], c1, c1, u0, u1);
        add2(tot.v[j], tot.v[j + 1], tot.v[j], tot.v[j + 1], u0, u1);
      }
      store8(dxt + eoff, tot);
      if (DROP) {  // gradient of the residual branch = mask * d x_new
        const f8 m = drop_mask8(dra, dk, row, r.c);
#pragma unroll
        for (int j = 0; j < 8; j += 2) mul2(tot.v[j], tot.v[j + 1], tot.v[j], tot.v[j + 1], m.v[j], m.v[j + 1]);
      }
      if (HAS_LN1) {
        const f8 aa = unpack8(a_raw), gg1 = unpack8(g1_raw);
        const float nm1 = -m1 * r1;
        float t1b = 0.f, t2b = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          fma2(xh.v[j], xh.v[j + 1], aa.v[j], aa.v[j + 1], r1, r1, nm1, nm1);
          fma2(acc[0].v[j], acc[0].v[j + 1], tot.v[j], tot.v[j + 1], xh.v[j], xh.v[j + 1], acc[0].v[j], acc[0].v[j + 1]);
          add2(acc[1].v[j], acc[1].v[j + 1], acc[1].v[j], acc[1].v[j + 1], tot.v[j], tot.v[j + 1]);
          mul2(d.v[j], d.v[j + 1], tot.v[j], tot.v[j + 1], gg1.v[j], gg1.v[j + 1]);
          fma2(t1, t1b, d.v[j], d.v[j + 1], xh.v[j], xh.v[j + 1], t1, t1b);
          add2(t2, t2b, t2, t2b, d.v[j], d.v[j + 1]);
        }
        t1 += t1b;
        t2 += t2b;
      } else {
        store8(da + eoff, tot);  // d a = d x_new
#pragma unroll
        for (int j = 0; j < 8; j += 2) add2(acc[4].v[j], acc[4].v[j + 1], acc[4].v[j], acc[4].v[j + 1], tot.v[j], tot.v[j + 1]);
      }
    }
    if (HAS_LN1) {
      group_sum2<NARROW>(t1, t2, red, flip, r, false);
      if (live) {
        const float c1 = -t1 * inv_n * r1, c2 = -t2 * inv_n * r1;
        f8 o;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          fma2(o.v[j], o.v[j + 1], d.v[j], d.v[j + 1], r1, r1, c2, c2);
          fma2(o.v[j], o.v[j + 1], xh.v[j], xh.v[j + 1], c1, c1, o.v[j], o.v[j + 1]);
          add2(acc[4].v[j], acc[4].v[j + 1], acc[4].v[j], acc[4].v[j + 1], o.v[j], o.v[j + 1]);
        }
        store8(da + eoff, o);
      }
    }
    if (++stage == NST) {
      stage = 0;
      phase ^= 1;
    }
  }
  flush_partials<5>(acc, partial, cols, r, reinterpret_cast<float*>(ring));
}

// ------------------------------------------------------------------------------------ colsum
// stage 1: grid (ceil(cols/64), CHUNKS); block (64, 4): partial[chunk, c] = sum over the chunk's rows
#define COLSUM_CHUNKS 64
template <typename T>
__global__ void colsum_stage1(const T* __restrict__ in, int64_t rows, int64_t cols, int64_t ld,
                              float* __restrict__ partial) {
  __shared__ float sm[4][64];
  pdl_launch();
  pdl_wait();
  const int64_t c = (int64_t)blockIdx.x * 64 + threadIdx.x;
  const int64_t per = (rows + COLSUM_CHUNKS - 1) / COLSUM_CHUNKS;
  const int64_t r0 = (int64_t)blockIdx.y * per;
  const int64_t r1 = min(rows, r0 + per);
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 4) s += (float)in[r * ld + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols)
    partial[(int64_t)blockIdx.y * cols + c] = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
}
template <typename TO>
__global__ void colsum_stage2(const float* __restrict__ partial, int64_t cols, TO* __restrict__ out, int accumulate) {
  pdl_launch();
  pdl_wait();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int k = 0; k < COLSUM_CHUNKS; ++k) s += partial[(int64_t)k * cols + c];
  if (accumulate) s += (float)out[c];
  out[c] = (TO)s;
}

}  // namespace

// out[s, c] = sum_r partial[s, r, c] for the OFAB_LN_PARTIAL_ROWS rows of every slab: one launch finishes
// dgamma / dbeta (/ dtype / dcls) of a backward kernel.  grid (ceil(cols/32), ns), block (32, 32).
template <typename TO>
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int cols, TO* __restrict__ out) {
  __shared__ float sm[32][33];
  pdl_launch();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  const float* base = partial + (int64_t)blockIdx.y * OFAB_LN_PARTIAL_ROWS * cols;
  float s = 0.f;
  if (c < cols) {
    constexpr int NR = (OFAB_LN_PARTIAL_ROWS + 31) / 32;
    float v[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {  // every load in flight before the first add (same order of adds as a plain loop)
      const int r = threadIdx.y + 32 * k;
      v[k] = r < OFAB_LN_PARTIAL_ROWS ? base[(int64_t)r * cols + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) s += v[k];
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    out[(int64_t)blockIdx.y * cols + c] = (TO)t;
  }
}

template <typename TO>
__global__ void reduce_rows_kernel(const float* __restrict__ partial, int64_t rows, int cols, TO* __restrict__ out) {
  __shared__ float sm[32][33];
  pdl_launch();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  const float* base = partial + (int64_t)blockIdx.y * rows * cols;
  float s = 0.f;
  if (c < cols)
    for (int64_t r0 = threadIdx.y; r0 < rows; r0 += 32 * 8) {  // eight loads in flight per thread
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int64_t r = r0 + 32 * k;
        v[k] = r < rows ? base[r * cols + c] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    out[(int64_t)blockIdx.y * cols + c] = (TO)t;
  }
}

// ---------------------------------------------------------------------------------- dispatch
struct LnLaunch {
  int tpr, rpb, block, grid;
};
// `target`: threads per block the kernel was compiled for at 2 CTAs / SM (384: <= 80 registers; 288: <= 112)
static inline LnLaunch ln_launch(int cols, int target = 384) {
  LnLaunch l;
  l.tpr = ((cols / 8) + 31) / 32 * 32;             // threads per row: one 8-column vector each
  l.rpb = l.tpr >= target ? 1 : target / l.tpr;    // rows per block iteration (wider rows: one row, up to 512 threads, 1 CTA / SM)
  l.block = l.rpb * l.tpr;
  l.grid = ofab_sm_count() * (l.block > target ? 1 : 2);  // persistent: resident CTAs (see LN_BOUNDS)
  if (l.grid > OFAB_LN_PARTIAL_ROWS) l.grid = OFAB_LN_PARTIAL_ROWS;
  return l;
}
// dynamic shared memory: mbarriers + `nst` stages of rpb rows of all staged inputs (`bytes_per_col` = sum of their
// element sizes); backward kernels reuse the ring to combine the row groups' partial sums (block * 8 floats)
static inline int ln_smem(const LnLaunch& l, int cols, int bytes_per_col, int nst, bool bwd) {
  const int ring = nst * l.rpb * cols * bytes_per_col;
  const int fb = bwd ? l.block * 8 * (int)sizeof(float) : 0;
  return kLnBarBytes + (ring > fb ? ring : fb);
}
constexpr int kLnMaxSmem = 200 * 1024;
static int ln_configure(const void* kern) {  // opt in to > 48 KB dynamic shared memory, once per kernel
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(kern)) return OFAB_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnMaxSmem);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "LayerNorm: cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  done.insert(kern);
  return OFAB_OK;
}
#define LN_ALIGNED16(p) ((((uintptr_t)(p)) & 15) == 0)
// optional dropout descriptor -> kernel form (`has` = descriptor given and not a no-op)
#define LN_DROP(who)                                                             \
  DropArgs da{};                                                                 \
  const bool has_drop = drop != nullptr && (drop->p > 0.f || drop->drop_path > 0.f); \
  if (has_drop && !ofab_drop_args(drop, da, who)) return OFAB_ERR_ARG
// launch kernel template K<..., 384> or K<..., 512> by block size
#define LN_GO(GRID, SMEM, TARGET, K384N, K384, K512, ...)                       \
  do {                                                                          \
    auto k384n = K384N; /* rows at most 4 warps wide */                         \
    auto k384 = K384;                                                           \
    auto k512 = K512;                                                           \
    const bool narrow_ = l.tpr <= 128;                                          \
    const void* kp = l.block > (TARGET) ? (const void*)k512 : (narrow_ ? (const void*)k384n : (const void*)k384); \
    int rc_ = ln_configure(kp);                                                 \
    if (rc_) return rc_;                                                        \
    if (l.block > (TARGET)) ofab_launch(k512, dim3(GRID), dim3(l.block), SMEM, st, __VA_ARGS__); \
    else if (narrow_) ofab_launch(k384n, dim3(GRID), dim3(l.block), SMEM, st, __VA_ARGS__); \
    else ofab_launch(k384, dim3(GRID), dim3(l.block), SMEM, st, __VA_ARGS__);   \
  } while (0)

extern "C" int ofab_ln_fwd(const void* x, int x_dt, const void* gamma, const void* beta, void* y, int y_dt, float* mean,
                           float* rstd, int64_t rows, int cols, float eps, int gelu, const ofab_dropout* drop, ofab_stream_t stream) {
  LN_DROP("ofab_ln_fwd");
  OFAB_REQUIRE(!has_drop || (gelu && x_dt == OFAB_BF16 && y_dt == OFAB_BF16), "ofab_ln_fwd: dropout is fused only into the bf16 GELU form");
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(rows >= 0, "ofab_ln_fwd: rows < 0");
  OFAB_REQUIRE(LN_ALIGNED16(x) && LN_ALIGNED16(y) && LN_ALIGNED16(gamma) && LN_ALIGNED16(beta), "ofab_ln_fwd: x / y / gamma / beta must be 16-byte aligned");
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // wide GELU rows (ffn_layernorm over 4d): two vectors per thread, half as many threads per row
  const bool two = gelu && x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && cols >= 2048 && cols % 16 == 0 && !getenv("OFAB_LN_VPT1");
  LnLaunch l = ln_launch(cols);
  if (two) {
    l.tpr = ((cols / 16) + 31) / 32 * 32;
    l.rpb = l.tpr >= 384 ? 1 : 384 / l.tpr;
    l.block = l.rpb * l.tpr;
    l.grid = ofab_sm_count() * 2;
    if (l.grid > OFAB_LN_PARTIAL_ROWS) l.grid = OFAB_LN_PARTIAL_ROWS;
  }
  const int64_t ngroups = (rows + l.rpb - 1) / l.rpb;
  const int grid = (int)(ngroups < l.grid ? ngroups : l.grid);
  const int smem = ln_smem(l, cols, x_dt == OFAB_F32 ? 4 : 2, 6, false);
#define ARGS(TX, TY) (const TX*)x, (const bf16*)gamma, (const bf16*)beta, (TY*)y, mean, rstd, rows, cols, eps, l.tpr, da
#define FWD(TX, TY, G) LN_GO(grid, smem, 384, (ln_fwd_kernel<TX, TY, G, 384, false, 1, true>), (ln_fwd_kernel<TX, TY, G, 384>), (ln_fwd_kernel<TX, TY, G, 512>), ARGS(TX, TY))
  if (two && has_drop) LN_GO(grid, smem, 384, (ln_fwd_kernel<bf16, bf16, true, 384, true, 2>), (ln_fwd_kernel<bf16, bf16, true, 384, true, 2>), (ln_fwd_kernel<bf16, bf16, true, 512, true>), ARGS(bf16, bf16));
  else if (two) LN_GO(grid, smem, 384, (ln_fwd_kernel<bf16, bf16, true, 384, false, 2>), (ln_fwd_kernel<bf16, bf16, true, 384, false, 2>), (ln_fwd_kernel<bf16, bf16, true, 512>), ARGS(bf16, bf16));
  else if (has_drop) LN_GO(grid, smem, 384, (ln_fwd_kernel<bf16, bf16, true, 384, true, 1, true>), (ln_fwd_kernel<bf16, bf16, true, 384, true>), (ln_fwd_kernel<bf16, bf16, true, 512, true>), ARGS(bf16, bf16));
  else if (x_dt == OFAB_F32 && y_dt == OFAB_BF16 && !gelu) FWD(float, bf16, false);
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && !gelu) FWD(bf16, bf16, false);
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && gelu) FWD(bf16, bf16, true);
  else if (x_dt == OFAB_F32 && y_dt == OFAB_F32 && !gelu) FWD(float, float, false);
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_F32 && !gelu) FWD(bf16, float, false);
  else {
    ofab_set_error("ofab_ln_fwd: unsupported dtype combination x=%d y=%d gelu=%d", x_dt, y_dt, gelu);
    return OFAB_ERR_ARG;
  }
#undef FWD
#undef ARGS
  OFAB_LAUNCH_CHECK("ofab_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const void* gamma, const float* mean,
                           const float* rstd, void* dx, int dx_dt, int dx_accum, float* dgb_partial, int64_t rows,
                           int cols, int gelu, const ofab_dropout* drop, ofab_stream_t stream) {
  LN_DROP("ofab_ln_bwd");
  OFAB_REQUIRE(!has_drop || (gelu && dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16), "ofab_ln_bwd: dropout is fused only into the bf16 GELU form");
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(!dx_accum || dx_dt == OFAB_F32, "ofab_ln_bwd: dx_accum needs fp32 dx");
  OFAB_REQUIRE(LN_ALIGNED16(x) && LN_ALIGNED16(dy) && LN_ALIGNED16(dx) && LN_ALIGNED16(gamma) && LN_ALIGNED16(dgb_partial),
               "ofab_ln_bwd: x / dy / dx / gamma / dgb_partial must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  const int grid = OFAB_LN_PARTIAL_ROWS;  // every partial row is written (idle CTAs write zeros)
  const int smem = ln_smem(l, cols, (x_dt == OFAB_F32 ? 4 : 2) + (dy_dt == OFAB_F32 ? 4 : 2), 4, true);
#define ARGS(TDY, TX, TDX) (const TDY*)dy, (const TX*)x, (const bf16*)gamma, mean, rstd, (TDX*)dx, dgb_partial, rows, cols, l.tpr, da
#define BWD(TDY, TX, TDX, G, A) LN_GO(grid, smem, 384, (ln_bwd_kernel<TDY, TX, TDX, G, A, 384, false, true>), (ln_bwd_kernel<TDY, TX, TDX, G, A, 384>), (ln_bwd_kernel<TDY, TX, TDX, G, A, 512>), ARGS(TDY, TX, TDX))
  if (has_drop) LN_GO(grid, smem, 384, (ln_bwd_kernel<bf16, bf16, bf16, true, false, 384, true, true>), (ln_bwd_kernel<bf16, bf16, bf16, true, false, 384, true>), (ln_bwd_kernel<bf16, bf16, bf16, true, false, 512, true>), ARGS(bf16, bf16, bf16));
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && dx_accum) BWD(bf16, float, float, false, true);
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum) BWD(bf16, float, float, false, false);
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && gelu) BWD(bf16, bf16, bf16, true, false);
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu) BWD(bf16, bf16, bf16, false, false);
  else if (dy_dt == OFAB_F32 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum) BWD(float, float, float, false, false);
  else if (dy_dt == OFAB_F32 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu) BWD(float, bf16, bf16, false, false);
  else {
    ofab_set_error("ofab_ln_bwd: unsupported dtype combination dy=%d x=%d dx=%d gelu=%d accum=%d", dy_dt, x_dt, dx_dt, gelu, dx_accum);
    return OFAB_ERR_ARG;
  }
#undef BWD
#undef ARGS
  OFAB_LAUNCH_CHECK("ofab_ln_bwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_fwd(const void* a, const float* x, const void* g1, const void* b1, const void* g2,
                                  const void* b2, float* x_new, void* y, float* stats, int64_t rows, int cols,
                                  float eps, const ofab_dropout* drop, ofab_stream_t stream) {
  LN_DROP("ofab_ln_res_ln_fwd");
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(LN_ALIGNED16(a) && LN_ALIGNED16(x) && LN_ALIGNED16(x_new) && LN_ALIGNED16(y), "ofab_ln_res_ln_fwd: a / x / x_new / y must be 16-byte aligned");
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  const int64_t ngroups = (rows + l.rpb - 1) / l.rpb;
  const int grid = (int)(ngroups < l.grid ? ngroups : l.grid);
  const int smem = ln_smem(l, cols, 6, 3, false);
#define RFWD(L1, D)                                                                                                                     \
  LN_GO(grid, smem, 384, (ln_res_ln_fwd_kernel<L1, 384, D, true>), (ln_res_ln_fwd_kernel<L1, 384, D>), (ln_res_ln_fwd_kernel<L1, 512, D>), (const bf16*)a, x, (const bf16*)g1, \
        (const bf16*)b1, (const bf16*)g2, (const bf16*)b2, x_new, (bf16*)y, stats, rows, cols, eps, l.tpr, da)
  if (g1 != nullptr) {
    if (has_drop) RFWD(true, true); else RFWD(true, false);
  } else {
    if (has_drop) RFWD(false, true); else RFWD(false, false);
  }
#undef RFWD
  OFAB_LAUNCH_CHECK("ofab_ln_res_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_bwd(const float* dx_new, const void* dy, const void* a, const float* x_new, const void* g1,
                                  const void* g2, const float* stats, float* dx_tot, void* da_out, float* dgb_partial,
                                  int64_t rows, int cols, const ofab_dropout* drop, ofab_stream_t stream) {
  void* da = da_out;
  DropArgs dra{};
  const bool has_drop = drop != nullptr && (drop->p > 0.f || drop->drop_path > 0.f);
  if (has_drop && !ofab_drop_args(drop, dra, "ofab_ln_res_ln_bwd")) return OFAB_ERR_ARG;
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(LN_ALIGNED16(dx_new) && LN_ALIGNED16(dy) && LN_ALIGNED16(x_new) && LN_ALIGNED16(dx_tot) && LN_ALIGNED16(da) &&
                   LN_ALIGNED16(dgb_partial) && (g1 == nullptr || LN_ALIGNED16(a)),
               "ofab_ln_res_ln_bwd: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = OFAB_LN_PARTIAL_ROWS;
  if (g1 != nullptr) {
    const LnLaunch l = ln_launch(cols, 288);  // 5 accumulator slabs: ~110 registers -> 288-thread CTAs, still two per SM
    const int smem = ln_smem(l, cols, 12, 2, true);
#define RBWD(D)                                                                                                                          \
  LN_GO(grid, smem, 288, (ln_res_ln_bwd_kernel<true, 288, D, true>), (ln_res_ln_bwd_kernel<true, 288, D>), (ln_res_ln_bwd_kernel<true, 512, D>), dx_new, (const bf16*)dy, (const bf16*)a, \
        x_new, (const bf16*)g1, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols, l.tpr, dra)
    if (has_drop) RBWD(true); else RBWD(false);
#undef RBWD
  } else {
    const LnLaunch l = ln_launch(cols, has_drop ? 288 : 384);  // the dropout form needs the mask registers: 288-thread CTAs
    const int smem = ln_smem(l, cols, 10, 2, true);
#define RBWD(D, T)                                                                                                                    \
  LN_GO(grid, smem, T, (ln_res_ln_bwd_kernel<false, T, D, true>), (ln_res_ln_bwd_kernel<false, T, D>), (ln_res_ln_bwd_kernel<false, 512, D>), dx_new, (const bf16*)dy, \
        (const bf16*)nullptr, x_new, (const bf16*)nullptr, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols, l.tpr, dra)
    if (has_drop) RBWD(true, 288); else RBWD(false, 384);
#undef RBWD
  }
  OFAB_LAUNCH_CHECK("ofab_ln_res_ln_bwd");
  return OFAB_OK;
}

// out[c] (+)= sum_r in[r, c]; deterministic two-stage reduction through caller scratch.
extern "C" int64_t ofab_colsum_scratch_elems(int64_t cols) { return (int64_t)COLSUM_CHUNKS * cols; }

extern "C" int ofab_colsum(const void* in, int in_dt, int64_t rows, int64_t cols, int64_t ld, void* out, int out_dt,
                           int accumulate, float* scratch, ofab_stream_t stream) {
  OFAB_REQUIRE(cols > 0 && rows >= 0 && ld >= cols, "ofab_colsum: bad shape rows=%lld cols=%lld ld=%lld", (long long)rows, (long long)cols, (long long)ld);
  OFAB_REQUIRE(scratch != nullptr, "ofab_colsum: scratch is NULL (need ofab_colsum_scratch_elems(cols) floats)");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((cols + 63) / 64), COLSUM_CHUNKS), block(64, 4);
  if (in_dt == OFAB_F32)
    ofab_launch((colsum_stage1<float>), dim3(grid), dim3(block), 0, st, (const float*)in, rows, cols, ld, scratch);
  else
    ofab_launch((colsum_stage1<bf16>), dim3(grid), dim3(block), 0, st, (const bf16*)in, rows, cols, ld, scratch);
  OFAB_LAUNCH_CHECK("ofab_colsum stage1");
  const int g2 = (int)((cols + 255) / 256);
  if (out_dt == OFAB_F32)
    ofab_launch((colsum_stage2<float>), dim3(g2), dim3(256), 0, st, scratch, cols, (float*)out, accumulate);
  else
    ofab_launch((colsum_stage2<bf16>), dim3(g2), dim3(256), 0, st, scratch, cols, (bf16*)out, accumulate);
  OFAB_LAUNCH_CHECK("ofab_colsum stage2");
  return OFAB_OK;
}

extern "C" int ofab_reduce_partials(const float* partial, int nslabs, int cols, void* out, int out_dt, ofab_stream_t stream) {
  OFAB_REQUIRE(nslabs > 0 && cols > 0, "ofab_reduce_partials: bad shape");
  dim3 grid((cols + 31) / 32, nslabs), block(32, 32);
  if (out_dt == OFAB_F32)
    ofab_launch((reduce_partials_kernel<float>), dim3(grid), dim3(block), 0, (cudaStream_t)stream, partial, cols, (float*)out);
  else
    ofab_launch((reduce_partials_kernel<bf16>), dim3(grid), dim3(block), 0, (cudaStream_t)stream, partial, cols, (bf16*)out);
  OFAB_LAUNCH_CHECK("ofab_reduce_partials");
  return OFAB_OK;
}

extern "C" int ofab_reduce_rows(const float* partial, int nslabs, int64_t rows, int cols, void* out, int out_dt, ofab_stream_t stream) {
  OFAB_REQUIRE(partial != nullptr && out != nullptr && nslabs > 0 && rows > 0 && cols > 0, "ofab_reduce_rows: bad arguments");
  dim3 grid((cols + 31) / 32, nslabs), block(32, 32);
  if (out_dt == OFAB_F32)
    ofab_launch((reduce_rows_kernel<float>), grid, block, 0, (cudaStream_t)stream, partial, rows, cols, (float*)out);
  else
    ofab_launch((reduce_rows_kernel<bf16>), grid, block, 0, (cudaStream_t)stream, partial, rows, cols, (bf16*)out);
  OFAB_LAUNCH_CHECK("ofab_reduce_rows");
  return OFAB_OK;
}
