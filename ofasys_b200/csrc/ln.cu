// LayerNorm family (HBM-bound): plain / GELU-fused LN, the fused LN -> +residual -> LN junction,
// and the column reductions that finish dgamma/dbeta and bias gradients.
//
// Layout: one row is owned by TPR threads (32 = one warp, or 128 = four warps); each thread keeps
// NV vectors of 8 consecutive columns in registers, so a row is read from HBM exactly once and
// written once.  Backward kernels are persistent over rows (fixed grid of OFAB_LN_PARTIAL_ROWS
// blocks) so per-column dgamma/dbeta partial sums stay in registers and leave as one row per block.
#include "common.cuh"

#define OFAB_LN_PARTIAL_ROWS 592  // 4 x 148 SMs

extern "C" int ofab_ln_partial_rows(void) { return OFAB_LN_PARTIAL_ROWS; }

namespace {

template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* red /* [4] per row-group */) {
  v = warp_sum(v);
  if (TPR == 32) return v;
  // TPR == 128: four warps cooperate through shared memory
  const int w = (threadIdx.x >> 5) & 3;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// ------------------------------------------------------------------------------------ forward
template <typename TX, typename TY, int TPR, int NV, bool GELU>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const TX* __restrict__ x, const bf16* __restrict__ gamma,
                                                     const bf16* __restrict__ beta, TY* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                     int64_t rows, int cols, float eps) {
  __shared__ float red[4];
  constexpr int RPB = 128 / TPR;
  const int lane = threadIdx.x % TPR;
  const int rib = threadIdx.x / TPR;
  const float inv_n = 1.0f / (float)cols;
  f8 g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * TPR) * 8;
    if (c < cols) {
      g[i] = load8(gamma + c);
      b[i] = load8(beta + c);
    }
  }
  const int64_t nblk_rows = (rows + RPB - 1) / RPB;
  for (int64_t rb = blockIdx.x; rb < nblk_rows; rb += gridDim.x) {
    const int64_t row = rb * RPB + rib;
    const bool live = row < rows;
    f8 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        v[i] = load8(x + row * cols + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (GELU) v[i].v[j] = gelu_f(v[i].v[j]);
          s += v[i].v[j];
        }
      }
    }
    const float mu = row_sum<TPR>(s, red) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i].v[j] - mu;
          q += d * d;
        }
      }
    }
    const float rs = rsqrtf(row_sum<TPR>(q, red) * inv_n + eps);
    if (live) {
      if (lane == 0) {
        mean[row] = mu;
        rstd[row] = rs;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (c < cols) {
          f8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - mu) * rs * g[i].v[j] + b[i].v[j];
          store8(y + row * cols + c, o);
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------- backward
// Reduce the per-thread column partials of the RPB rows in a block and write one partial row.
template <int TPR, int NV, int NS>
__device__ __forceinline__ void flush_partials(f8 (&acc)[NS][NV], float* __restrict__ partial, int cols) {
  constexpr int RPB = 128 / TPR;
  const int lane = threadIdx.x % TPR;
  const int rib = threadIdx.x / TPR;
  if (RPB == 1) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (c < cols)
          store8(partial + ((int64_t)s * OFAB_LN_PARTIAL_ROWS + blockIdx.x) * cols + c, acc[s][i]);
      }
  } else {
    // TPR == 32, NV == 1 here (cols <= 256): 4 rows per block share columns
    __shared__ float buf[RPB][TPR * 8 * NV + 8];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) buf[rib][(lane + i * TPR) * 8 + j] = acc[s][i].v[j];
      __syncthreads();
      for (int c = threadIdx.x; c < cols; c += 128) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < RPB; ++r) t += buf[r][c];
        partial[((int64_t)s * OFAB_LN_PARTIAL_ROWS + blockIdx.x) * cols + c] = t;
      }
    }
  }
}

template <typename TDY, typename TX, typename TDX, int TPR, int NV, bool GELU, bool ACCUM>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                     const bf16* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, TDX* __restrict__ dx,
                                                     float* __restrict__ partial, int64_t rows, int cols) {
  __shared__ float red[4];
  constexpr int RPB = 128 / TPR;
  const int lane = threadIdx.x % TPR;
  const int rib = threadIdx.x / TPR;
  const float inv_n = 1.0f / (float)cols;
  f8 acc[3][NV];  // dgamma, dbeta, column sums of dx (= bias gradient of the Linear that produced x)
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][i].v[j] = acc[1][i].v[j] = acc[2][i].v[j] = 0.f;
  const int64_t nblk_rows = (rows + RPB - 1) / RPB;
  for (int64_t rb = blockIdx.x; rb < nblk_rows; rb += gridDim.x) {
    const int64_t row = rb * RPB + rib;
    const bool live = row < rows;
    const float mu = live ? mean[row] : 0.f;
    const float rs = live ? rstd[row] : 0.f;
    f8 xh[NV], d[NV], pre[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        pre[i] = load8(x + row * cols + c);
        d[i] = load8(dy + row * cols + c);
        const f8 g = load8(gamma + c);  // L1-resident; not kept in registers
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = pre[i].v[j];
          if (GELU) {  // one erf evaluation serves both gelu(x) and gelu'(x); `pre` keeps the derivative from here on
            float cdf, px;
            gelu_parts(a, cdf, px);
            pre[i].v[j] = cdf + px;
            a *= cdf;
          }
          xh[i].v[j] = (a - mu) * rs;
          acc[0][i].v[j] += d[i].v[j] * xh[i].v[j];
          acc[1][i].v[j] += d[i].v[j];
          const float gg = d[i].v[j] * g.v[j];
          d[i].v[j] = gg;
          s1 += gg * xh[i].v[j];
          s2 += gg;
        }
      }
    }
    const float c1 = row_sum<TPR>(s1, red) * inv_n;
    const float c2 = row_sum<TPR>(s2, red) * inv_n;
    if (live) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (c < cols) {
          f8 o;
          if (ACCUM) o = load8(reinterpret_cast<const TDX*>(dx) + row * cols + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float t = rs * (d[i].v[j] - c2 - xh[i].v[j] * c1);
            if (GELU) t *= pre[i].v[j];
            acc[2][i].v[j] += t;
            o.v[j] = ACCUM ? o.v[j] + t : t;
          }
          store8(dx + row * cols + c, o);
        }
      }
    }
  }
  flush_partials<TPR, NV, 3>(acc, partial, cols);
}

// ---------------------------------------------------------------- fused LN -> +res -> LN
// HAS_LN1 = false: x_new = x + a (no first LayerNorm): the deferred residual add of an FFN output fused
// with the next block's pre-LayerNorm.
template <int TPR, int NV, bool HAS_LN1>
__global__ void __launch_bounds__(128) ln_res_ln_fwd_kernel(const bf16* __restrict__ a, const float* __restrict__ x,
                                                            const bf16* __restrict__ g1, const bf16* __restrict__ b1,
                                                            const bf16* __restrict__ g2, const bf16* __restrict__ b2,
                                                            float* __restrict__ x_new, bf16* __restrict__ y,
                                                            float* __restrict__ stats, int64_t rows, int cols, float eps) {
  __shared__ float red[4];
  constexpr int RPB = 128 / TPR;
  const int lane = threadIdx.x % TPR;
  const int rib = threadIdx.x / TPR;
  const float inv_n = 1.0f / (float)cols;
  const int64_t nblk_rows = (rows + RPB - 1) / RPB;
  for (int64_t rb = blockIdx.x; rb < nblk_rows; rb += gridDim.x) {
    const int64_t row = rb * RPB + rib;
    const bool live = row < rows;
    f8 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        v[i] = load8(a + row * cols + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i].v[j];
      }
    }
    float m1 = 0.f, r1 = 1.f, q = 0.f;
    if (HAS_LN1) {
      m1 = row_sum<TPR>(s, red) * inv_n;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (live && c < cols) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = v[i].v[j] - m1;
            q += d * d;
          }
        }
      }
      r1 = rsqrtf(row_sum<TPR>(q, red) * inv_n + eps);
    }
    // x_new = x + LN1(a)   (or x + a)
    s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        const f8 xx = load8(x + row * cols + c);
        if (HAS_LN1) {
          const f8 gg = load8(g1 + c), bb = load8(b1 + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i].v[j] = xx.v[j] + ((v[i].v[j] - m1) * r1 * gg.v[j] + bb.v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i].v[j] += xx.v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i].v[j];
        store8(x_new + row * cols + c, v[i]);
      }
    }
    const float m2 = row_sum<TPR>(s, red) * inv_n;
    q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i].v[j] - m2;
          q += d * d;
        }
      }
    }
    const float r2 = rsqrtf(row_sum<TPR>(q, red) * inv_n + eps);
    if (live) {
      if (lane == 0) {
        stats[row] = m1;
        stats[rows + row] = r1;
        stats[2 * rows + row] = m2;
        stats[3 * rows + row] = r2;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (c < cols) {
          const f8 gg = load8(g2 + c), bb = load8(b2 + c);
          f8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = (v[i].v[j] - m2) * r2 * gg.v[j] + bb.v[j];
          store8(y + row * cols + c, o);
        }
      }
    }
  }
}

template <int TPR, int NV, bool HAS_LN1>
__global__ void __launch_bounds__(128) ln_res_ln_bwd_kernel(const float* __restrict__ dxn, const bf16* __restrict__ dy,
                                                            const bf16* __restrict__ a, const float* __restrict__ x_new,
                                                            const bf16* __restrict__ g1, const bf16* __restrict__ g2,
                                                            const float* __restrict__ stats, float* __restrict__ dxt,
                                                            bf16* __restrict__ da, float* __restrict__ partial,
                                                            int64_t rows, int cols) {
  __shared__ float red[4];
  constexpr int RPB = 128 / TPR;
  const int lane = threadIdx.x % TPR;
  const int rib = threadIdx.x / TPR;
  const float inv_n = 1.0f / (float)cols;
  f8 acc[5][NV];  // dg1, db1, dg2, db2, column sums of da (bias gradient of the Linear that produced a)
#pragma unroll
  for (int s = 0; s < 5; ++s)
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[s][i].v[j] = 0.f;
  const int64_t nblk_rows = (rows + RPB - 1) / RPB;
  for (int64_t rb = blockIdx.x; rb < nblk_rows; rb += gridDim.x) {
    const int64_t row = rb * RPB + rib;
    const bool live = row < rows;
    const float m1 = live ? stats[row] : 0.f, r1 = live ? stats[rows + row] : 0.f;
    const float m2 = live ? stats[2 * rows + row] : 0.f, r2 = live ? stats[3 * rows + row] : 0.f;
    f8 xh[NV], d[NV];
    float s1 = 0.f, s2 = 0.f;
    // ---- LN2 backward on x_new
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        const f8 xx = load8(x_new + row * cols + c), gg = load8(g2 + c);
        d[i] = load8(dy + row * cols + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i].v[j] = (xx.v[j] - m2) * r2;
          acc[2][i].v[j] += d[i].v[j] * xh[i].v[j];
          acc[3][i].v[j] += d[i].v[j];
          const float t = d[i].v[j] * gg.v[j];
          d[i].v[j] = t;
          s1 += t * xh[i].v[j];
          s2 += t;
        }
      }
    }
    float c1 = row_sum<TPR>(s1, red) * inv_n;
    float c2 = row_sum<TPR>(s2, red) * inv_n;
    s1 = s2 = 0.f;
    // ---- dx_tot = dx_new + LN2'(dy); LN1 backward on a with upstream dx_tot
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * TPR) * 8;
      if (live && c < cols) {
        const f8 up = load8(dxn + row * cols + c);
        f8 tot;
#pragma unroll
        for (int j = 0; j < 8; ++j) tot.v[j] = up.v[j] + r2 * (d[i].v[j] - c2 - xh[i].v[j] * c1);
        store8(dxt + row * cols + c, tot);
        if (HAS_LN1) {
          const f8 aa = load8(a + row * cols + c), gg = load8(g1 + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xh[i].v[j] = (aa.v[j] - m1) * r1;
            acc[0][i].v[j] += tot.v[j] * xh[i].v[j];
            acc[1][i].v[j] += tot.v[j];
            const float t = tot.v[j] * gg.v[j];
            d[i].v[j] = t;
            s1 += t * xh[i].v[j];
            s2 += t;
          }
        } else {
          store8(da + row * cols + c, tot);  // d a = d x_new
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[4][i].v[j] += tot.v[j];
        }
      }
    }
    if (HAS_LN1) {
      c1 = row_sum<TPR>(s1, red) * inv_n;
      c2 = row_sum<TPR>(s2, red) * inv_n;
    }
    if (live && HAS_LN1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (lane + i * TPR) * 8;
        if (c < cols) {
          f8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o.v[j] = r1 * (d[i].v[j] - c2 - xh[i].v[j] * c1);
            acc[4][i].v[j] += o.v[j];
          }
          store8(da + row * cols + c, o);
        }
      }
    }
  }
  flush_partials<TPR, NV, 5>(acc, partial, cols);
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
#define LN_SHAPE_DISPATCH(cols, CALL)                 \
  if ((cols) <= 256) { CALL(32, 1); }                 \
  else if ((cols) <= 1024) { CALL(128, 1); }          \
  else if ((cols) <= 2048) { CALL(128, 2); }          \
  else if ((cols) <= 3072) { CALL(128, 3); }          \
  else { CALL(128, 4); }

static inline int ln_fwd_grid(int64_t rows, int tpr) {
  const int rpb = 128 / tpr;
  int64_t nb = (rows + rpb - 1) / rpb;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  return (int)(nb < cap ? (nb > 0 ? nb : 1) : cap);
}

extern "C" int ofab_ln_fwd(const void* x, int x_dt, const void* gamma, const void* beta, void* y, int y_dt,
                           float* mean, float* rstd, int64_t rows, int cols, float eps, int gelu,
                           ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(rows >= 0, "ofab_ln_fwd: rows < 0");
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TPR, NV)                                                                                        \
  {                                                                                                          \
    const int grid = ln_fwd_grid(rows, TPR);                                                                 \
    if (x_dt == OFAB_F32 && y_dt == OFAB_BF16 && !gelu)                                                      \
      ln_fwd_kernel<float, bf16, TPR, NV, false><<<grid, 128, 0, st>>>((const float*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, mean, rstd, rows, cols, eps); \
    else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && !gelu)                                                \
      ln_fwd_kernel<bf16, bf16, TPR, NV, false><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, mean, rstd, rows, cols, eps); \
    else if (x_dt == OFAB_BF16 && y_dt == OFAB_BF16 && gelu)                                                 \
      ln_fwd_kernel<bf16, bf16, TPR, NV, true><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, mean, rstd, rows, cols, eps); \
    else if (x_dt == OFAB_F32 && y_dt == OFAB_F32 && !gelu)                                                  \
      ln_fwd_kernel<float, float, TPR, NV, false><<<grid, 128, 0, st>>>((const float*)x, (const bf16*)gamma, (const bf16*)beta, (float*)y, mean, rstd, rows, cols, eps); \
    else if (x_dt == OFAB_BF16 && y_dt == OFAB_F32 && !gelu)                                                 \
      ln_fwd_kernel<bf16, float, TPR, NV, false><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, (float*)y, mean, rstd, rows, cols, eps); \
    else {                                                                                                   \
      ofab_set_error("ofab_ln_fwd: unsupported dtype combination x=%d y=%d gelu=%d", x_dt, y_dt, gelu);      \
      return OFAB_ERR_ARG;                                                                                   \
    }                                                                                                        \
  }
  LN_SHAPE_DISPATCH(cols, CALL)
#undef CALL
  OFAB_LAUNCH_CHECK("ofab_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const void* gamma, const float* mean,
                           const float* rstd, void* dx, int dx_dt, int dx_accum, float* dgb_partial, int64_t rows,
                           int cols, int gelu, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  OFAB_REQUIRE(!dx_accum || dx_dt == OFAB_F32, "ofab_ln_bwd: dx_accum needs fp32 dx");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = OFAB_LN_PARTIAL_ROWS;
#define CALL(TPR, NV)                                                                                        \
  {                                                                                                          \
    if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && dx_accum)                    \
      ln_bwd_kernel<bf16, float, float, TPR, NV, false, true><<<grid, 128, 0, st>>>((const bf16*)dy, (const float*)x, (const bf16*)gamma, mean, rstd, (float*)dx, dgb_partial, rows, cols); \
    else if (dy_dt == OFAB_BF16 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum)              \
      ln_bwd_kernel<bf16, float, float, TPR, NV, false, false><<<grid, 128, 0, st>>>((const bf16*)dy, (const float*)x, (const bf16*)gamma, mean, rstd, (float*)dx, dgb_partial, rows, cols); \
    else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && gelu)                          \
      ln_bwd_kernel<bf16, bf16, bf16, TPR, NV, true, false><<<grid, 128, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)gamma, mean, rstd, (bf16*)dx, dgb_partial, rows, cols); \
    else if (dy_dt == OFAB_BF16 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu)                         \
      ln_bwd_kernel<bf16, bf16, bf16, TPR, NV, false, false><<<grid, 128, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)gamma, mean, rstd, (bf16*)dx, dgb_partial, rows, cols); \
    else if (dy_dt == OFAB_F32 && x_dt == OFAB_F32 && dx_dt == OFAB_F32 && !gelu && !dx_accum)               \
      ln_bwd_kernel<float, float, float, TPR, NV, false, false><<<grid, 128, 0, st>>>((const float*)dy, (const float*)x, (const bf16*)gamma, mean, rstd, (float*)dx, dgb_partial, rows, cols); \
    else if (dy_dt == OFAB_F32 && x_dt == OFAB_BF16 && dx_dt == OFAB_BF16 && !gelu)                          \
      ln_bwd_kernel<float, bf16, bf16, TPR, NV, false, false><<<grid, 128, 0, st>>>((const float*)dy, (const bf16*)x, (const bf16*)gamma, mean, rstd, (bf16*)dx, dgb_partial, rows, cols); \
    else {                                                                                                   \
      ofab_set_error("ofab_ln_bwd: unsupported dtype combination dy=%d x=%d dx=%d gelu=%d accum=%d", dy_dt, x_dt, dx_dt, gelu, dx_accum); \
      return OFAB_ERR_ARG;                                                                                   \
    }                                                                                                        \
  }
  LN_SHAPE_DISPATCH(cols, CALL)
#undef CALL
  OFAB_LAUNCH_CHECK("ofab_ln_bwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_fwd(const void* a, const float* x, const void* g1, const void* b1, const void* g2,
                                  const void* b2, float* x_new, void* y, float* stats, int64_t rows, int cols,
                                  float eps, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_fwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  if (rows == 0) return OFAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TPR, NV)                                                                                       \
  if (g1 != nullptr)                                                                                        \
    ln_res_ln_fwd_kernel<TPR, NV, true><<<ln_fwd_grid(rows, TPR), 128, 0, st>>>((const bf16*)a, x, (const bf16*)g1, (const bf16*)b1, (const bf16*)g2, (const bf16*)b2, x_new, (bf16*)y, stats, rows, cols, eps); \
  else                                                                                                      \
    ln_res_ln_fwd_kernel<TPR, NV, false><<<ln_fwd_grid(rows, TPR), 128, 0, st>>>((const bf16*)a, x, nullptr, nullptr, (const bf16*)g2, (const bf16*)b2, x_new, (bf16*)y, stats, rows, cols, eps);
  LN_SHAPE_DISPATCH(cols, CALL)
#undef CALL
  OFAB_LAUNCH_CHECK("ofab_ln_res_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ln_res_ln_bwd(const float* dx_new, const void* dy, const void* a, const float* x_new, const void* g1,
                                  const void* g2, const float* stats, float* dx_tot, void* da, float* dgb_partial,
                                  int64_t rows, int cols, ofab_stream_t stream) {
  OFAB_REQUIRE(cols % 8 == 0 && cols >= 8 && cols <= 4096, "ofab_ln_res_ln_bwd: cols=%d must be a multiple of 8 in [8,4096]", cols);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(TPR, NV)                                                                                       \
  if (g1 != nullptr)                                                                                        \
    ln_res_ln_bwd_kernel<TPR, NV, true><<<OFAB_LN_PARTIAL_ROWS, 128, 0, st>>>(dx_new, (const bf16*)dy, (const bf16*)a, x_new, (const bf16*)g1, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols); \
  else                                                                                                      \
    ln_res_ln_bwd_kernel<TPR, NV, false><<<OFAB_LN_PARTIAL_ROWS, 128, 0, st>>>(dx_new, (const bf16*)dy, nullptr, x_new, nullptr, (const bf16*)g2, stats, dx_tot, (bf16*)da, dgb_partial, rows, cols);
  LN_SHAPE_DISPATCH(cols, CALL)
#undef CALL
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
