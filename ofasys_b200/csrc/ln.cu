// LayerNorm family (HBM-bound): plain / GELU-fused LN, the fused LN -> +residual -> LN junction,
// and the column reductions that finish dgamma/dbeta and bias gradients.
//
// Layout: one row is owned by cols/8 threads (rounded up to whole warps), one 8-column vector per thread, so a
// row is read from HBM exactly once and written once and register use stays low enough for ~36 resident warps
// per SM.  Backward kernels are persistent over rows (fixed grid of OFAB_LN_PARTIAL_ROWS blocks) so per-column
// dgamma/dbeta partial sums stay in registers and leave as one row per block.
#include "common.cuh"

#define OFAB_LN_PARTIAL_ROWS 888  // 6 x 148 SMs: persistent backward grid

extern "C" int ofab_ln_partial_rows(void) { return OFAB_LN_PARTIAL_ROWS; }

namespace {

// Thread layout of every LayerNorm kernel: a row is owned by `tpr` threads (a multiple of 32), each holding ONE
// vector of 8 consecutive columns, so per-thread state is tiny and many warps stay resident to cover HBM
// latency; a block of 384 (512 for cols > 3072) threads works on blockDim/tpr rows at a time.
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
// sum over the threads of one row group; all row groups of the block call this together
__device__ __forceinline__ float group_sum(float v, float* red /* [16][16] */, const RowCtx& r) {
  v = warp_sum(v);
  if (r.wpr == 1) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && r.rib < 16) red[r.rib * 16 + r.wir] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < r.wpr; ++w) s += red[(r.rib < 16 ? r.rib : 0) * 16 + w];
  return s;
}
__device__ __forceinline__ void group_sum2(float& a, float& b, float* red /* [2][16][16] */, const RowCtx& r) {
  a = warp_sum(a);
  b = warp_sum(b);
  if (r.wpr == 1) return;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && r.rib < 16) {
    red[r.rib * 16 + r.wir] = a;
    red[256 + r.rib * 16 + r.wir] = b;
  }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
  for (int w = 0; w < r.wpr; ++w) {
    sa += red[(r.rib < 16 ? r.rib : 0) * 16 + w];
    sb += red[256 + (r.rib < 16 ? r.rib : 0) * 16 + w];
  }
  a = sa;
  b = sb;
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

// ------------------------------------------------------------------------------------ forward
template <typename TX, typename TY, bool GELU>
__global__ void __launch_bounds__(512) ln_fwd_kernel(const TX* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                                     TY* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows,
                                                     int cols, float eps, int tpr) {
  __shared__ float red[256];
  const RowCtx r = row_ctx(tpr, cols);
  const float inv_n = 1.0f / (float)cols;
  f8 g, b;
  if (r.col_ok) {
    g = load8(gamma + r.c);
    b = load8(beta + r.c);
  }
  const int64_t ngroups = (rows + r.rpb - 1) / r.rpb;
  for (int64_t rb = blockIdx.x; rb < ngroups; rb += gridDim.x) {
    const int64_t row = rb * r.rpb + r.rib;
    const bool live = r.col_ok && row < rows;
    f8 v;
    float s = 0.f;
    if (live) {
      v = load8(x + row * cols + r.c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (GELU) v.v[j] = gelu_f(v.v[j]);
        s += v.v[j];
      }
    }
    const float mu = group_sum(s, red, r) * inv_n;
    float q = 0.f;
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v.v[j] - mu;
        q += d * d;
      }
    }
    const float rs = rsqrtf(group_sum(q, red, r) * inv_n + eps);
    if (live) {
      if (r.lane_in_row == 0) {
        mean[row] = mu;
        rstd[row] = rs;
      }
      f8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (v.v[j] - mu) * rs * g.v[j] + b.v[j];
      store8(y + row * cols + r.c, o);
    }
  }
}

// ----------------------------------------------------------------------------------- backward
template <typename TDY, typename TX, typename TDX, bool GELU, bool ACCUM>
__global__ void __launch_bounds__(512) ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const bf16* __restrict__ gamma,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd, TDX* __restrict__ dx,
                                                     float* __restrict__ partial, int64_t rows, int cols, int tpr) {
  __shared__ float red[512];
  extern __shared__ float fbuf[];
  const RowCtx r = row_ctx(tpr, cols);
  const float inv_n = 1.0f / (float)cols;
  f8 g;
  if (r.col_ok) g = load8(gamma + r.c);
  f8 acc[3];  // dgamma, dbeta, column sums of dx (= bias gradient of the Linear that produced x)
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s].v[j] = 0.f;
  const int64_t ngroups = (rows + r.rpb - 1) / r.rpb;
  for (int64_t rb = blockIdx.x; rb < ngroups; rb += gridDim.x) {
    const int64_t row = rb * r.rpb + r.rib;
    const bool live = r.col_ok && row < rows;
    float mu = 0.f, rs = 0.f;
    f8 xh, d, gp;
    float s1 = 0.f, s2 = 0.f;
    if (live) {
      mu = mean[row];
      rs = rstd[row];
      const f8 pre = load8(x + row * cols + r.c);
      d = load8(dy + row * cols + r.c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = pre.v[j];
        if (GELU) {  // one erf evaluation serves both gelu(x) and gelu'(x)
          float cdf, px;
          gelu_parts(a, cdf, px);
          gp.v[j] = cdf + px;
          a *= cdf;
        }
        xh.v[j] = (a - mu) * rs;
        acc[0].v[j] += d.v[j] * xh.v[j];
        acc[1].v[j] += d.v[j];
        const float gg = d.v[j] * g.v[j];
        d.v[j] = gg;
        s1 += gg * xh.v[j];
        s2 += gg;
      }
    }
    group_sum2(s1, s2, red, r);
    const float c1 = s1 * inv_n, c2 = s2 * inv_n;
    if (live) {
      f8 o;
      if (ACCUM) o = load8(reinterpret_cast<const TDX*>(dx) + row * cols + r.c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = rs * (d.v[j] - c2 - xh.v[j] * c1);
        if (GELU) t *= gp.v[j];
        acc[2].v[j] += t;
        o.v[j] = ACCUM ? o.v[j] + t : t;
      }
      store8(dx + row * cols + r.c, o);
    }
  }
  flush_partials<3>(acc, partial, cols, r, fbuf);
}

// ---------------------------------------------------------------- fused LN -> +res -> LN
// HAS_LN1 = false: x_new = x + a (no first LayerNorm): the deferred residual add of an FFN output fused
// with the next block's pre-LayerNorm.
template <bool HAS_LN1>
__global__ void __launch_bounds__(512) ln_res_ln_fwd_kernel(const bf16* __restrict__ a, const float* __restrict__ x, const bf16* __restrict__ g1,
                                                            const bf16* __restrict__ b1, const bf16* __restrict__ g2, const bf16* __restrict__ b2,
                                                            float* __restrict__ x_new, bf16* __restrict__ y, float* __restrict__ stats,
                                                            int64_t rows, int cols, float eps, int tpr) {
  __shared__ float red[256];
  const RowCtx r = row_ctx(tpr, cols);
  const float inv_n = 1.0f / (float)cols;
  f8 gg1, bb1, gg2, bb2;
  if (r.col_ok) {
    if (HAS_LN1) {
      gg1 = load8(g1 + r.c);
      bb1 = load8(b1 + r.c);
    }
    gg2 = load8(g2 + r.c);
    bb2 = load8(b2 + r.c);
  }
  const int64_t ngroups = (rows + r.rpb - 1) / r.rpb;
  for (int64_t rb = blockIdx.x; rb < ngroups; rb += gridDim.x) {
    const int64_t row = rb * r.rpb + r.rib;
    const bool live = r.col_ok && row < rows;
    f8 v, xx;
    float s = 0.f;
    if (live) {
      v = load8(a + row * cols + r.c);
      xx = load8(x + row * cols + r.c);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v.v[j];
    }
    float m1 = 0.f, r1 = 1.f;
    if (HAS_LN1) {
      m1 = group_sum(s, red, r) * inv_n;
      float q = 0.f;
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v.v[j] - m1;
          q += d * d;
        }
      }
      r1 = rsqrtf(group_sum(q, red, r) * inv_n + eps);
    }
    s = 0.f;
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v.v[j] = HAS_LN1 ? xx.v[j] + ((v.v[j] - m1) * r1 * gg1.v[j] + bb1.v[j]) : xx.v[j] + v.v[j];
        s += v.v[j];
      }
      store8(x_new + row * cols + r.c, v);
    }
    const float m2 = group_sum(s, red, r) * inv_n;
    float q = 0.f;
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v.v[j] - m2;
        q += d * d;
      }
    }
    const float r2 = rsqrtf(group_sum(q, red, r) * inv_n + eps);
    if (live) {
      if (r.lane_in_row == 0) {
        stats[row] = m1;
        stats[rows + row] = r1;
        stats[2 * rows + row] = m2;
        stats[3 * rows + row] = r2;
      }
      f8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (v.v[j] - m2) * r2 * gg2.v[j] + bb2.v[j];
      store8(y + row * cols + r.c, o);
    }
  }
}

template <bool HAS_LN1>
__global__ void __launch_bounds__(512) ln_res_ln_bwd_kernel(const float* __restrict__ dxn, const bf16* __restrict__ dy, const bf16* __restrict__ a,
                                                            const float* __restrict__ x_new, const bf16* __restrict__ g1,
                                                            const bf16* __restrict__ g2, const float* __restrict__ stats,
                                                            float* __restrict__ dxt, bf16* __restrict__ da, float* __restrict__ partial,
                                                            int64_t rows, int cols, int tpr) {
  __shared__ float red[512];
  extern __shared__ float fbuf[];
  const RowCtx r = row_ctx(tpr, cols);
  const float inv_n = 1.0f / (float)cols;
  f8 gg1, gg2;
  if (r.col_ok) {
    if (HAS_LN1) gg1 = load8(g1 + r.c);
    gg2 = load8(g2 + r.c);
  }
  f8 acc[5];  // dg1, db1, dg2, db2, column sums of da (bias gradient of the Linear that produced a)
#pragma unroll
  for (int s = 0; s < 5; ++s)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s].v[j] = 0.f;
  const int64_t ngroups = (rows + r.rpb - 1) / r.rpb;
  for (int64_t rb = blockIdx.x; rb < ngroups; rb += gridDim.x) {
    const int64_t row = rb * r.rpb + r.rib;
    const bool live = r.col_ok && row < rows;
    float m1 = 0.f, r1 = 0.f, m2 = 0.f, r2 = 0.f;
    f8 xh, d, up, aa;
    float s1 = 0.f, s2 = 0.f;
    if (live) {
      m1 = stats[row]; r1 = stats[rows + row]; m2 = stats[2 * rows + row]; r2 = stats[3 * rows + row];
      const f8 xx = load8(x_new + row * cols + r.c);
      d = load8(dy + row * cols + r.c);
      up = load8(dxn + row * cols + r.c);
      if (HAS_LN1) aa = load8(a + row * cols + r.c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh.v[j] = (xx.v[j] - m2) * r2;
        acc[2].v[j] += d.v[j] * xh.v[j];
        acc[3].v[j] += d.v[j];
        const float t = d.v[j] * gg2.v[j];
        d.v[j] = t;
        s1 += t * xh.v[j];
        s2 += t;
      }
    }
    group_sum2(s1, s2, red, r);
    float c1 = s1 * inv_n, c2 = s2 * inv_n;
    s1 = s2 = 0.f;
    f8 tot;
    if (live) {
#pragma unroll
      for (int j = 0; j < 8; ++j) tot.v[j] = up.v[j] + r2 * (d.v[j] - c2 - xh.v[j] * c1);
      store8(dxt + row * cols + r.c, tot);
      if (HAS_LN1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh.v[j] = (aa.v[j] - m1) * r1;
          acc[0].v[j] += tot.v[j] * xh.v[j];
          acc[1].v[j] += tot.v[j];
          const float t = tot.v[j] * gg1.v[j];
          d.v[j] = t;
          s1 += t * xh.v[j];
          s2 += t;
        }
      } else {
        store8(da + row * cols + r.c, tot);  // d a = d x_new
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[4].v[j] += tot.v[j];
      }
    }
    if (HAS_LN1) {
      group_sum2(s1, s2, red, r);
      c1 = s1 * inv_n;
      c2 = s2 * inv_n;
      if (live) {
        f8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o.v[j] = r1 * (d.v[j] - c2 - xh.v[j] * c1);
          acc[4].v[j] += o.v[j];
        }
        store8(da + row * cols + r.c, o);
      }
    }
  }
  flush_partials<5>(acc, partial, cols, r, fbuf);
}

// ------------------------------------------------------------------------------------ colsum
// stage 1: grid (ceil(cols/64), CHUNKS); block (64, 4): partial[chunk, c] = sum over the chunk's rows
#define COLSUM_CHUNKS 64
template <typename T>
__global__ void colsum_stage1(const T* __restrict__ in, int64_t rows, int64_t cols, int64_t ld,
                              float* __restrict__ partial) {
  __shared__ float sm[4][64];
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
  const int c = blockIdx.x * 32 + threadIdx.x;
  const float* base = partial + (int64_t)blockIdx.y * OFAB_LN_PARTIAL_ROWS * cols;
  float s = 0.f;
  if (c < cols)
    for (int r = threadIdx.y; r < OFAB_LN_PARTIAL_ROWS; r += 32) s += base[(int64_t)r * cols + c];
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
  int tpr, block, smem;
};
static inline LnLaunch ln_launch(int cols) {
  LnLaunch l;
  l.tpr = ((cols / 8) + 31) / 32 * 32;           // threads per row: one 8-column vector each
  l.block = l.tpr >= 96 ? l.tpr : 384;           // one row per block (3..16 warps); narrow rows share a block
  l.smem = l.block * 8 * (int)sizeof(float);     // row-group combine buffer of the backward kernels
  return l;
}
static inline int ln_fwd_grid(int64_t rows, const LnLaunch& l) {
  const int rpb = l.block / l.tpr;
  int64_t nb = (rows + rpb - 1) / rpb;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  return (int)(nb < cap ? (nb > 0 ? nb : 1) : cap);
}

extern "C" int ofab_ln_fwd(const void* x, int x_dt, const void* gamma, const void* beta, void* y, int y_dt, float* mean,
                           float* rstd, int64_t rows, int cols, float eps, int gelu, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(rows >= 0, "ofab_ln_fwd: rows < 0");
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  const int grid = ln_fwd_grid(rows, l);
#define ARGS(TX, TY) (const TX*)x, (const bf16*)gamma, (const bf16*)beta, (TY*)y, mean, rstd, rows, cols, eps, l.tpr
  if (x_dt == OFAB_F32 && y_dt == OFAB_BF16 && !gelu) ln_fwd_kernel<float, bf16, false><<<grid, l.block, 0, st>>>(ARGS(float, bf16));
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && !gelu) ln_fwd_kernel<bf16, bf16, false><<<grid, l.block, 0, st>>>(ARGS(bf16, bf16));
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && gelu) ln_fwd_kernel<bf16, bf16, true><<<grid, l.block, 0, st>>>(ARGS(bf16, bf16));
  else if (x_dt == OFAB_F32 && y_dt == OFAB_F32 && !gelu) ln_fwd_kernel<float, float, false><<<grid, l.block, 0, st>>>(ARGS(float, float));
  else if (x_dt == OFAB_BF16 && y_dt == OFAB_F32 && !gelu) ln_fwd_kernel<bf16, float, false><<<grid, l.block, 0, st>>>(ARGS(bf16, float));
  else {
    ofab_set_error("ofab_ln_fwd: unsupported dtype combination x=%d y=%d gelu=%d", x_dt, y_dt, gelu);
    return OFAB_ERR_ARG;
  }
#undef ARGS
  OFAB_LAUNCH_CHECK("ofab_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const void* gamma, const float* mean,
                           const float* rstd, void* dx, int dx_dt, int dx_accum, float* dgb_partial, int64_t rows,
                           int cols, int gelu, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(!dx_accum || dx_dt == OFAB_F32, "ofab_ln_bwd: dx_accum needs fp32 dx");
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  const int grid = OFAB_LN_PARTIAL_ROWS;
#define ARGS(TDY, TX, TDX) (const TDY*)dy, (const TX*)x, (const bf16*)gamma, mean, rstd, (TDX*)dx, dgb_partial, rows, cols, l.tpr
  if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && dx_accum)
    ln_bwd_kernel<bf16, float, float, false, true><<<grid, l.block, l.smem, st>>>(ARGS(bf16, float, float));
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum)
    ln_bwd_kernel<bf16, float, float, false, false><<<grid, l.block, l.smem, st>>>(ARGS(bf16, float, float));
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && gelu)
    ln_bwd_kernel<bf16, bf16, bf16, true, false><<<grid, l.block, l.smem, st>>>(ARGS(bf16, bf16, bf16));
  else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu)
    ln_bwd_kernel<bf16, bf16, bf16, false, false><<<grid, l.block, l.smem, st>>>(ARGS(bf16, bf16, bf16));
  else if (dy_dt == OFAB_F32 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum)
    ln_bwd_kernel<float, float, float, false, false><<<grid, l.block, l.smem, st>>>(ARGS(float, float, float));
  else if (dy_dt == OFAB_F32 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu)
    ln_bwd_kernel<float, bf16, bf16, false, false><<<grid, l.block, l.smem, st>>>(ARGS(float, bf16, bf16));
  else {
    ofab_set_error("ofab_ln_bwd: unsupported dtype combination dy=%d x=%d dx=%d gelu=%d accum=%d", dy_dt, x_dt, dx_dt, gelu, dx_accum);
    return OFAB_ERR_ARG;
  }
#undef ARGS
  OFAB_LAUNCH_CHECK("ofab_ln_bwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_fwd(const void* a, const float* x, const void* g1, const void* b1, const void* g2,
                                  const void* b2, float* x_new, void* y, float* stats, int64_t rows, int cols,
                                  float eps, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  const int grid = ln_fwd_grid(rows, l);
  if (g1 != nullptr)
    ln_res_ln_fwd_kernel<true><<<grid, l.block, 0, st>>>((const bf16*)a, x, (const bf16*)g1, (const bf16*)b1, (const bf16*)g2, (const bf16*)b2, x_new, (bf16*)y, stats, rows, cols, eps, l.tpr);
  else
    ln_res_ln_fwd_kernel<false><<<grid, l.block, 0, st>>>((const bf16*)a, x, nullptr, nullptr, (const bf16*)g2, (const bf16*)b2, x_new, (bf16*)y, stats, rows, cols, eps, l.tpr);
  OFAB_LAUNCH_CHECK("ofab_ln_res_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_bwd(const float* dx_new, const void* dy, const void* a, const float* x_new, const void* g1,
                                  const void* g2, const float* stats, float* dx_tot, void* da, float* dgb_partial,
                                  int64_t rows, int cols, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  cudaStream_t st = (cudaStream_t)stream;
  const LnLaunch l = ln_launch(cols);
  if (g1 != nullptr)
    ln_res_ln_bwd_kernel<true><<<OFAB_LN_PARTIAL_ROWS, l.block, l.smem, st>>>(dx_new, (const bf16*)dy, (const bf16*)a, x_new, (const bf16*)g1, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols, l.tpr);
  else
    ln_res_ln_bwd_kernel<false><<<OFAB_LN_PARTIAL_ROWS, l.block, l.smem, st>>>(dx_new, (const bf16*)dy, nullptr, x_new, nullptr, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols, l.tpr);
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
    colsum_stage1<float><<<grid, block, 0, st>>>((const float*)in, rows, cols, ld, scratch);
  else
    colsum_stage1<bf16><<<grid, block, 0, st>>>((const bf16*)in, rows, cols, ld, scratch);
  OFAB_LAUNCH_CHECK("ofab_colsum stage1");
  const int g2 = (int)((cols + 255) / 256);
  if (out_dt == OFAB_F32)
    colsum_stage2<float><<<g2, 256, 0, st>>>(scratch, cols, (float*)out, accumulate);
  else
    colsum_stage2<bf16><<<g2, 256, 0, st>>>(scratch, cols, (bf16*)out, accumulate);
  OFAB_LAUNCH_CHECK("ofab_colsum stage2");
  return OFAB_OK;
}

extern "C" int ofab_reduce_partials(const float* partial, int nslabs, int cols, void* out, int out_dt, ofab_stream_t stream) {
  OFAB_REQUIRE(nslabs > 0 && cols > 0, "ofab_reduce_partials: bad shape");
  dim3 grid((cols + 31) / 32, nslabs), block(32, 32);
  if (out_dt == OFAB_F32)
    reduce_partials_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(partial, cols, (float*)out);
  else
    reduce_partials_kernel<bf16><<<grid, block, 0, (cudaStream_t)stream>>>(partial, cols, (bf16*)out);
  OFAB_LAUNCH_CHECK("ofab_reduce_partials");
  return OFAB_OK;
}
