// ResNet backbone pieces of the image / video adaptors (ofasys/module/resnet.py:116-246) around the tcgen05
// GEMM.  Activations are channel-last bf16 [B, H, W, C] (= token rows [B*H*W, C]), so 1x1 convolutions are plain
// GEMMs, 3x3 / 7x7 convolutions are im2col (vector copies) + GEMM, and BatchNorm (training mode: batch statistics,
// resnet.py keeps nn.BatchNorm2d in train mode unless freeze_resnet) is a column reduction + an elementwise pass
// fused with the residual add and ReLU of the bottleneck.
#include "common.cuh"

namespace {

inline int ew_grid(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- im2col from an NCHW image (stem conv 7x7 s2 p3): cols[(b,ho,wo), c*kh*kw + i*kw + j] (weight's own flattening)
template <typename T>
__global__ void im2col_nchw_kernel(const T* __restrict__ img, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                                   bf16* __restrict__ cols, int64_t ldk) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int kk = C * k * k;
  const int64_t total = (int64_t)B * Ho * Wo * ldk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ldk;
    const int col = (int)(i % ldk);
    float v = 0.f;
    if (col < kk) {
      const int c = col / (k * k), ki = (col / k) % k, kj = col % k;
      const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho), b = (int)(row / ((int64_t)Wo * Ho));
      const int y = ho * stride - pad + ki, x = wo * stride - pad + kj;
      if (y >= 0 && y < H && x >= 0 && x < W) v = (float)img[(((int64_t)b * C + c) * H + y) * W + x];
    }
    cols[i] = __float2bfloat16(v);
  }
}

// ---- im2col on channel-last activations, square kernel k, stride, zero padding:
// cols[(b,ho,wo), (i*k + j)*C + c] = x[b, ho*stride - pad + i, wo*stride - pad + j, c]
__global__ void im2col_nhwc_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                                   bf16* __restrict__ cols) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8, kk = k * k;
  const int64_t total = (int64_t)B * Ho * Wo * kk * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int t = (int)((i / C8) % kk);
    const int64_t r = i / ((int64_t)C8 * kk);
    const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((int64_t)Wo * Ho));
    const int y = ho * stride - pad + t / k, xx = wo * stride - pad + t % k;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H && xx >= 0 && xx < W) v = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + y) * W + xx) * C + c8 * 8);
    *reinterpret_cast<uint4*>(cols + (r * kk + t) * C + c8 * 8) = v;
  }
}
// adjoint (gather form, no atomics): dx[b,y,x,c] = sum over taps (i,j) and outputs (ho,wo) with ho*stride - pad + i == y ...
__global__ void col2im_nhwc_kernel(const bf16* __restrict__ dcols, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                                   bf16* __restrict__ dx) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8, kk = k * k;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t r = i / C8;
    const int xx = (int)(r % W), y = (int)((r / W) % H), b = (int)(r / ((int64_t)W * H));
    f8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int ti = 0; ti < k; ++ti) {
      const int yy = y + pad - ti;
      if (yy < 0 || yy % stride != 0 || yy / stride >= Ho) continue;
      for (int tj = 0; tj < k; ++tj) {
        const int xw = xx + pad - tj;
        if (xw < 0 || xw % stride != 0 || xw / stride >= Wo) continue;
        const f8 v = load8(dcols + ((((int64_t)b * Ho + yy / stride) * Wo + xw / stride) * kk + ti * k + tj) * C + c8 * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j];
      }
    }
    store8(dx + r * C + c8 * 8, acc);
  }
}

// ---- spatial stride-2 subsampling (1x1 stride-2 downsample convs) and its adjoint
__global__ void subsample2_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, bf16* __restrict__ y, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t r = i / C8;
    const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((int64_t)Wo * Ho));
    *reinterpret_cast<uint4*>(y + r * C + c8 * 8) = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + 2 * ho) * W + 2 * wo) * C + c8 * 8);
  }
}
__global__ void subsample2_bwd_kernel(const bf16* __restrict__ dy, int B, int H, int W, int C, bf16* __restrict__ dx, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = (int64_t)B * H * W * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t r = i / C8;
    const int xx = (int)(r % W), y = (int)((r / W) % H), b = (int)(r / ((int64_t)W * H));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(y & 1) && !(xx & 1) && y / 2 < Ho && xx / 2 < Wo) v = *reinterpret_cast<const uint4*>(dy + (((int64_t)b * Ho + y / 2) * Wo + xx / 2) * C + c8 * 8);
    *reinterpret_cast<uint4*>(dx + r * C + c8 * 8) = v;
  }
}

// ---- max pooling 3x3 stride 2 pad 1 (channel-last); argmax tap (0..8, first maximum as torch) saved for backward
__global__ void maxpool_fwd_kernel(const bf16* __restrict__ x, int B, int H, int W, int C, bf16* __restrict__ y, uint8_t* __restrict__ arg, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int64_t total = (int64_t)B * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((int64_t)Wo * Ho));
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int y = 2 * ho - 1 + t / 3, xx = 2 * wo - 1 + t % 3;
      if (y >= 0 && y < H && xx >= 0 && xx < W) {
        const float v = __bfloat162float(x[(((int64_t)b * H + y) * W + xx) * C + c]);
        if (v > best) { best = v; bi = t; }
      }
    }
    y[i] = __float2bfloat16(best);
    arg[i] = (uint8_t)bi;
  }
}
__global__ void maxpool_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ arg, int B, int H, int W, int C, bf16* __restrict__ dx, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int xx = (int)(r % W), y = (int)((r / W) % H), b = (int)(r / ((int64_t)W * H));
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {  // output (ho, wo) whose tap t lands on (y, xx)
      const int yy = y + 1 - t / 3, xw = xx + 1 - t % 3;
      if (yy < 0 || (yy & 1) || yy / 2 >= Ho || xw < 0 || (xw & 1) || xw / 2 >= Wo) continue;
      const int64_t o = ((((int64_t)b * Ho + yy / 2) * Wo + xw / 2) * C) + c;
      if (arg[o] == t) s += __bfloat162float(dy[o]);
    }
    dx[i] = __float2bfloat16(s);
  }
}

// ---- BatchNorm (training): per-channel statistics over the R = B*H*W rows of x [R, C]
// Thread layout of the per-channel reductions (statistics, backward sums): a block of 256 threads is `tx` threads across
// the channels (one 16-byte vector of 8 channels each, tx = largest power of two <= min(C/8, 32)) by ty = 256 / tx rows;
// grid = (channel blocks, row chunks) with as many row chunks as it takes to fill the machine whatever C is (the 64-channel
// layers have ONE channel block).  Every block leaves one partial row per sum; stage 2 reduces the chunks 32 at a time.
struct BnRed {
  int tx, ty, gx, nch;
};
inline int bn_tx(int C) {
  const int c8 = C / 8;
  int tx = 1;
  while (tx * 2 <= c8 && tx < 32) tx *= 2;
  return tx;
}
inline int bn_max_chunks(int C) {
  const int tx = bn_tx(C), gx = (C / 8 + tx - 1) / tx;
  return (2 * ofab_sm_count() + gx - 1) / gx;
}
inline BnRed bn_red(int C, int64_t R) {
  BnRed r;
  r.tx = bn_tx(C);
  r.ty = 256 / r.tx;
  r.gx = (C / 8 + r.tx - 1) / r.tx;
  const int64_t by_rows = R / (2 * r.ty);  // at least two rows per thread
  const int64_t nch = std::min<int64_t>(bn_max_chunks(C), by_rows < 1 ? 1 : by_rows);
  r.nch = (int)nch;
  return r;
}
// block-level finish of a two-sum reduction: a, q of every thread -> one value per channel of the block and sum
// (var != NULL: the second sum is scaled by rstd of its channel)
__device__ __forceinline__ void bn_block_finish(const f8& a, const f8& q, const float* __restrict__ var, float eps, int C, float* __restrict__ part,
                                                int nch) {
  __shared__ float red[2][256 * 8];
  const int tx = blockDim.x, ty = blockDim.y;
  const int t = threadIdx.y * tx + threadIdx.x;
  store8(&red[0][t * 8], a);
  store8(&red[1][t * 8], q);
  __syncthreads();
  const int wch = tx * 8;  // channels of this block
  if (t < wch) {
    const int c = blockIdx.x * wch + t;
    if (c < C) {
      float sa = 0.f, sq = 0.f;
      for (int y = 0; y < ty; ++y) {
        sa += red[0][y * wch + t];
        sq += red[1][y * wch + t];
      }
      if (var != nullptr) sq *= rsqrtf(var[c] + eps);  // backward: sum g (x - mean) -> sum g xhat
      part[(int64_t)blockIdx.y * C + c] = sa;
      part[((int64_t)nch + blockIdx.y) * C + c] = sq;
    }
  }
}
// reduce the chunks of both sums for 32 channels: block (32 channels, 32 chunk lanes)
__device__ __forceinline__ void bn_chunk_sums(const float* __restrict__ part, int nch, int C, int c, float& a, float& q) {
  __shared__ float ra[32][33], rq[32][33];
  float sa = 0.f, sq = 0.f;
  if (c < C)
    for (int k = threadIdx.y; k < nch; k += 32) {
      sa += part[(int64_t)k * C + c];
      sq += part[((int64_t)nch + k) * C + c];
    }
  ra[threadIdx.y][threadIdx.x] = sa;
  rq[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  a = 0.f;
  q = 0.f;
  if (threadIdx.y == 0) {
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      a += ra[k][threadIdx.x];
      q += rq[k][threadIdx.x];
    }
  }
}

// stage 1: partial (sum, sum of squares) of (x - shift[c]) per row chunk; shift = first row (conditioning)
__global__ void __launch_bounds__(256) bn_stats_stage1(const bf16* __restrict__ x, int64_t R, int C, float* __restrict__ part /* [2][nch][C] */) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = c8 < C / 8;
  const int nch = gridDim.y;
  const int64_t per = (R + nch - 1) / nch;
  const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(R, r0 + per);
  f8 a, q;
#pragma unroll
  for (int j = 0; j < 8; ++j) a.v[j] = q.v[j] = 0.f;
  if (ok) {
    const f8 sh = load8(x + c8 * 8);
    const bf16* xp = x + c8 * 8;
    const int ty = blockDim.y;
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.y; r < r1; r += ty) {
      const f8 v = load8(xp + r * C);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v.v[j] - sh.v[j];
        a.v[j] += d;
        q.v[j] = fmaf(d, d, q.v[j]);
      }
    }
  }
  bn_block_finish(a, q, nullptr, 0.f, C, part, nch);
}
// stage 2: mean / biased variance; momentum update of the running statistics (unbiased variance), as nn.BatchNorm2d
template <typename TR>
__global__ void __launch_bounds__(1024) bn_stats_stage2(const bf16* __restrict__ x, const float* __restrict__ part, int nch, int64_t R, int C,
                                                        float* __restrict__ mean, float* __restrict__ var, TR* __restrict__ run_mean,
                                                        TR* __restrict__ run_var, float momentum) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a, q;
  bn_chunk_sums(part, nch, C, c, a, q);
  if (threadIdx.y != 0 || c >= C) return;
  const float sh = __bfloat162float(x[c]);
  const float m = a / (float)R;
  const float v = fmaxf(q / (float)R - m * m, 0.f);
  mean[c] = m + sh;
  var[c] = v;
  if (run_mean != nullptr) {
    const float unb = R > 1 ? v * (float)R / (float)(R - 1) : v;
    run_mean[c] = (TR)((1.f - momentum) * (float)run_mean[c] + momentum * (m + sh));
    run_var[c] = (TR)((1.f - momentum) * (float)run_var[c] + momentum * unb);
  }
}
// y = relu?( (x - mean) * rstd * gamma + beta (+ residual) ).  The grid stride is a multiple of C / 8 whenever C / 8 divides the
// block size (every ResNet width), so a thread keeps ITS 8 channels for the whole loop and the per-channel constants stay
// in registers; other widths reload them per vector.
__global__ void __launch_bounds__(256) bn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                                       const bf16* __restrict__ gamma, const bf16* __restrict__ beta, const bf16* __restrict__ res,
                                                       bf16* __restrict__ y, int64_t R, int C, float eps, int relu) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = R * C8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool fixed = stride % C8 == 0;
  f8 mu, sc, sf;
  auto consts = [&](int c) {
    const f8 g = load8(gamma + c);
    sf = load8(beta + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu.v[j] = mean[c + j];
      sc.v[j] = rsqrtf(var[c + j] + eps) * g.v[j];
    }
  };
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 < total) consts((int)(i0 % C8) * 8);
#pragma unroll 2
  for (int64_t i = i0; i < total; i += stride) {
    if (!fixed) consts((int)(i % C8) * 8);
    f8 v = load8(x + i * 8);
    f8 rr;
    if (res != nullptr) rr = load8(res + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(v.v[j] - mu.v[j], sc.v[j], sf.v[j]);
      if (res != nullptr) o += rr.v[j];
      v.v[j] = relu ? fmaxf(o, 0.f) : o;
    }
    store8(y + i * 8, v);
  }
}
// backward stage 1: per-channel partial sums of g and g*xhat, g = dy * (y > 0 if relu)
__global__ void __launch_bounds__(256) bn_bwd_stage1(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                                                     const float* __restrict__ mean, const float* __restrict__ var, int64_t R, int C, float eps,
                                                     int relu, float* __restrict__ part) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = c8 < C / 8;
  const int nch = gridDim.y;
  const int64_t per = (R + nch - 1) / nch;
  const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(R, r0 + per);
  f8 a, q;
#pragma unroll
  for (int j = 0; j < 8; ++j) a.v[j] = q.v[j] = 0.f;
  if (ok) {
    f8 m;
#pragma unroll
    for (int j = 0; j < 8; ++j) m.v[j] = mean[c8 * 8 + j];
    const int ty = blockDim.y;
#pragma unroll 2
    for (int64_t r = r0 + threadIdx.y; r < r1; r += ty) {
      const int64_t o = r * C + c8 * 8;
      f8 g = load8(dy + o);
      const f8 xv = load8(x + o);
      if (relu) {
        const f8 yv = load8(y + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] = yv.v[j] > 0.f ? g.v[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a.v[j] += g.v[j];
        q.v[j] = fmaf(g.v[j], xv.v[j] - m.v[j], q.v[j]);  // times rstd in the block finish
      }
    }
  }
  bn_block_finish(a, q, var, eps, C, part, nch);
}
__global__ void __launch_bounds__(1024) bn_bwd_stage2(const float* __restrict__ part, int nch, int C, float* __restrict__ sums /* [2][C]: sum g, sum g*xhat */) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a, q;
  bn_chunk_sums(part, nch, C, c, a, q);
  if (threadIdx.y != 0 || c >= C) return;
  sums[c] = a;
  sums[C + c] = q;
}
// dx = gamma * rstd * (g - sum_g/R - xhat * sum_gxhat/R);  dres = g.  Per-channel constants in registers as in bn_apply_kernel.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ y,
                                                           const float* __restrict__ mean, const float* __restrict__ var, const bf16* __restrict__ gamma,
                                                           const float* __restrict__ sums, bf16* __restrict__ dx, bf16* __restrict__ dres, int64_t R, int C,
                                                           float eps, int relu, int frozen_stats) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = R * C8;
  const float invR = 1.0f / (float)R;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool fixed = stride % C8 == 0;
  f8 mu, ka, kb, kc;  // dx = ka * (g - kb - (x - mu) * kc):  ka = gamma rstd, kb = sum_g / R, kc = rstd sum_gxhat / R
  auto consts = [&](int c) {
    const f8 gm = load8(gamma + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float rs = rsqrtf(var[c + j] + eps);
      mu.v[j] = mean[c + j];
      ka.v[j] = gm.v[j] * rs;
      // frozen_stats (eval-mode BatchNorm: mean / var are the running buffers, constants of the graph): dx = gamma * rstd * g
      kb.v[j] = frozen_stats ? 0.f : sums[c + j] * invR;
      kc.v[j] = frozen_stats ? 0.f : rs * sums[C + c + j] * invR;
    }
  };
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 < total) consts((int)(i0 % C8) * 8);
#pragma unroll 2
  for (int64_t i = i0; i < total; i += stride) {
    if (!fixed) consts((int)(i % C8) * 8);
    f8 g = load8(dy + i * 8);
    f8 xv;
    if (!frozen_stats) xv = load8(x + i * 8);
    if (relu) {
      const f8 yv = load8(y + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) g.v[j] = yv.v[j] > 0.f ? g.v[j] : 0.f;
    }
    if (dres != nullptr) store8(dres + i * 8, g);
    f8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o.v[j] = frozen_stats ? ka.v[j] * g.v[j] : ka.v[j] * fmaf(-(xv.v[j] - mu.v[j]), kc.v[j], g.v[j] - kb.v[j]);
    store8(dx + i * 8, o);
  }
}

}  // namespace

// ---- video clip [B, C, F, H, W] -> frames bf16 [B*F, C, H, W] + "frame is all zero" flags (the reference's frame
// padding test `clip.abs().mean() == 0`, video_image_sequence.py:136-139).  One block per frame.
namespace {
template <typename T>
__global__ void video_frames_kernel(const T* __restrict__ vid, int C, int F, int64_t HW, bf16* __restrict__ frames, uint8_t* __restrict__ zero) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int b = blockIdx.x / F, f = blockIdx.x % F;
  int any = 0;
  for (int c = 0; c < C; ++c) {
    const T* src = vid + (((int64_t)b * C + c) * F + f) * HW;
    bf16* dst = frames ? frames + (((int64_t)blockIdx.x) * C + c) * HW : nullptr;
    for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) {
      const float v = (float)src[i];
      any |= (v != 0.f);
      if (dst) dst[i] = __float2bfloat16(v);
    }
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && zero) zero[blockIdx.x] = any ? 0 : 1;
}
}  // namespace
extern "C" int ofab_video_frames(const void* video, int dt, int B, int C, int F, int64_t HW, void* frames, uint8_t* zero, ofab_stream_t stream) {
  OFAB_REQUIRE(B > 0 && C > 0 && F > 0 && HW > 0, "ofab_video_frames: empty clip");
  OFAB_REQUIRE(dt == OFAB_F32 || dt == OFAB_BF16, "ofab_video_frames: dtype must be f32 or bf16");
  if (dt == OFAB_F32)
    ofab_launch(video_frames_kernel<float>, dim3(B * F), dim3(256), (size_t)(0), (cudaStream_t)stream, (const float*)video, C, F, HW, (bf16*)frames, zero);
  else
    ofab_launch(video_frames_kernel<bf16>, dim3(B * F), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)video, C, F, HW, (bf16*)frames, zero);
  OFAB_LAUNCH_CHECK("ofab_video_frames");
  return OFAB_OK;
}
extern "C" int ofab_im2col_nchw(const void* img, int img_dt, int B, int C, int H, int W, int k, int stride, int pad, void* cols, int64_t ldk,
                                ofab_stream_t stream) {
  OFAB_REQUIRE(k > 0 && stride > 0 && pad >= 0 && ldk >= (int64_t)C * k * k && ldk % 8 == 0, "ofab_im2col_nchw: bad geometry");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)B * Ho * Wo * ldk;
  if (img_dt == OFAB_F32)
    ofab_launch(im2col_nchw_kernel<float>, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const float*)img, B, C, H, W, k, stride, pad, Ho, Wo, (bf16*)cols, ldk);
  else
    ofab_launch(im2col_nchw_kernel<bf16>, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)img, B, C, H, W, k, stride, pad, Ho, Wo, (bf16*)cols, ldk);
  OFAB_LAUNCH_CHECK("ofab_im2col_nchw");
  return OFAB_OK;
}
extern "C" int ofab_im2col_nhwc(const void* x, int B, int H, int W, int C, int k, int stride, int pad, void* cols, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0 && k > 0 && stride > 0 && pad >= 0, "ofab_im2col_nhwc: bad geometry (C %% 8 == 0)");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)B * Ho * Wo * k * k * (C / 8);
  ofab_launch(im2col_nhwc_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, B, H, W, C, k, stride, pad, Ho, Wo, (bf16*)cols);
  OFAB_LAUNCH_CHECK("ofab_im2col_nhwc");
  return OFAB_OK;
}
extern "C" int ofab_col2im_nhwc(const void* dcols, int B, int H, int W, int C, int k, int stride, int pad, void* dx, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0 && k > 0 && stride > 0 && pad >= 0, "ofab_col2im_nhwc: bad geometry (C %% 8 == 0)");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)B * H * W * (C / 8);
  ofab_launch(col2im_nhwc_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)dcols, B, H, W, C, k, stride, pad, Ho, Wo, (bf16*)dx);
  OFAB_LAUNCH_CHECK("ofab_col2im_nhwc");
  return OFAB_OK;
}
extern "C" int ofab_subsample2(const void* x, int B, int H, int W, int C, void* y, int backward, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0, "ofab_subsample2: C %% 8 != 0");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if (!backward) {
    const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
    ofab_launch(subsample2_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, B, H, W, C, (bf16*)y, Ho, Wo);
  } else {  // x = dy [B,Ho,Wo,C], y = dx [B,H,W,C]
    const int64_t total = (int64_t)B * H * W * (C / 8);
    ofab_launch(subsample2_bwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, B, H, W, C, (bf16*)y, Ho, Wo);
  }
  OFAB_LAUNCH_CHECK("ofab_subsample2");
  return OFAB_OK;
}
extern "C" int ofab_maxpool3x3s2_fwd(const void* x, int B, int H, int W, int C, void* y, uint8_t* argmax, ofab_stream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * C;
  ofab_launch(maxpool_fwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, B, H, W, C, (bf16*)y, argmax, Ho, Wo);
  OFAB_LAUNCH_CHECK("ofab_maxpool3x3s2_fwd");
  return OFAB_OK;
}
extern "C" int ofab_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, int B, int H, int W, int C, void* dx, ofab_stream_t stream) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * H * W * C;
  ofab_launch(maxpool_bwd_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)dy, argmax, B, H, W, C, (bf16*)dx, Ho, Wo);
  OFAB_LAUNCH_CHECK("ofab_maxpool3x3s2_bwd");
  return OFAB_OK;
}
extern "C" int64_t ofab_bn_scratch_elems(int C) { return (int64_t)2 * bn_max_chunks(C) * C; }

extern "C" int ofab_bn_stats(const void* x, int64_t R, int C, float* mean, float* var, void* run_mean, void* run_var, int run_dt,
                             float momentum, float* scratch, ofab_stream_t stream) {
  OFAB_REQUIRE(R > 0 && C > 0 && C % 8 == 0 && scratch != nullptr, "ofab_bn_stats: bad arguments (C %% 8 == 0)");
  cudaStream_t st = (cudaStream_t)stream;
  const BnRed rd = bn_red(C, R);
  dim3 grid(rd.gx, rd.nch), block(rd.tx, rd.ty);
  ofab_launch(bn_stats_stage1, dim3(grid), dim3(block), (size_t)(0), st, (const bf16*)x, R, C, scratch);
  OFAB_LAUNCH_CHECK("ofab_bn_stats stage1");
  if (run_dt == OFAB_F32)
    ofab_launch(bn_stats_stage2<float>, dim3((C + 31) / 32), dim3(32, 32), (size_t)(0), st, (const bf16*)x, (const float*)scratch, rd.nch, R, C, mean, var, (float*)run_mean, (float*)run_var, momentum);
  else
    ofab_launch(bn_stats_stage2<bf16>, dim3((C + 31) / 32), dim3(32, 32), (size_t)(0), st, (const bf16*)x, (const float*)scratch, rd.nch, R, C, mean, var, (bf16*)run_mean, (bf16*)run_var, momentum);
  OFAB_LAUNCH_CHECK("ofab_bn_stats stage2");
  return OFAB_OK;
}
extern "C" int ofab_bn_apply(const void* x, const float* mean, const float* var, const void* gamma, const void* beta, const void* residual,
                             void* y, int64_t R, int C, float eps, int relu, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0, "ofab_bn_apply: C %% 8 != 0");
  ofab_launch(bn_apply_kernel, dim3(ew_grid(R * (C / 8), 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, mean, var, (const bf16*)gamma, (const bf16*)beta,
                                                                              (const bf16*)residual, (bf16*)y, R, C, eps, relu);
  OFAB_LAUNCH_CHECK("ofab_bn_apply");
  return OFAB_OK;
}
extern "C" int ofab_bn_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* var, const void* gamma, float* sums,
                           void* dx, void* dres, int64_t R, int C, float eps, int relu, float* scratch, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0 && scratch != nullptr && sums != nullptr, "ofab_bn_bwd: bad arguments");
  OFAB_REQUIRE(!relu || y != nullptr, "ofab_bn_bwd: relu needs the forward output");
  cudaStream_t st = (cudaStream_t)stream;
  const BnRed rd = bn_red(C, R);
  dim3 grid(rd.gx, rd.nch), block(rd.tx, rd.ty);
  ofab_launch(bn_bwd_stage1, dim3(grid), dim3(block), (size_t)(0), st, (const bf16*)dy, (const bf16*)x, (const bf16*)y, mean, var, R, C, eps, relu, scratch);
  OFAB_LAUNCH_CHECK("ofab_bn_bwd stage1");
  ofab_launch(bn_bwd_stage2, dim3((C + 31) / 32), dim3(32, 32), (size_t)(0), st, (const float*)scratch, rd.nch, C, sums);
  OFAB_LAUNCH_CHECK("ofab_bn_bwd stage2");
  ofab_launch(bn_bwd_apply_kernel, dim3(ew_grid(R * (C / 8), 256)), dim3(256), (size_t)(0), st, (const bf16*)dy, (const bf16*)x, (const bf16*)y, mean, var, (const bf16*)gamma, sums,
                                                                (bf16*)dx, (bf16*)dres, R, C, eps, relu, 0);
  OFAB_LAUNCH_CHECK("ofab_bn_bwd apply");
  return OFAB_OK;
}

// eval-mode (frozen) BatchNorm backward: y = relu?((x - running_mean) * rsqrt(running_var + eps) * gamma + beta (+ residual))
extern "C" int ofab_bn_bwd_eval(const void* dy, const void* x, const void* y, const float* mean, const float* var, const void* gamma, float* sums,
                                void* dx, void* dres, int64_t R, int C, float eps, int relu, float* scratch, ofab_stream_t stream) {
  OFAB_REQUIRE(C % 8 == 0 && dx != nullptr, "ofab_bn_bwd_eval: bad arguments");
  OFAB_REQUIRE(!relu || y != nullptr, "ofab_bn_bwd_eval: relu needs the forward output");
  OFAB_REQUIRE(sums == nullptr || scratch != nullptr, "ofab_bn_bwd_eval: parameter gradients need scratch");
  cudaStream_t st = (cudaStream_t)stream;
  if (sums != nullptr) {  // dbeta / dgamma wanted (BatchNorm affine parameters not frozen)
    const BnRed rd = bn_red(C, R);
    dim3 grid(rd.gx, rd.nch), block(rd.tx, rd.ty);
    ofab_launch(bn_bwd_stage1, dim3(grid), dim3(block), (size_t)(0), st, (const bf16*)dy, (const bf16*)x, (const bf16*)y, mean, var, R, C, eps, relu, scratch);
    OFAB_LAUNCH_CHECK("ofab_bn_bwd_eval stage1");
    ofab_launch(bn_bwd_stage2, dim3((C + 31) / 32), dim3(32, 32), (size_t)(0), st, (const float*)scratch, rd.nch, C, sums);
    OFAB_LAUNCH_CHECK("ofab_bn_bwd_eval stage2");
  }
  ofab_launch(bn_bwd_apply_kernel, dim3(ew_grid(R * (C / 8), 256)), dim3(256), (size_t)(0), st, (const bf16*)dy, (const bf16*)x, (const bf16*)y, mean, var, (const bf16*)gamma, sums,
                                                                (bf16*)dx, (bf16*)dres, R, C, eps, relu, 1);
  OFAB_LAUNCH_CHECK("ofab_bn_bwd_eval apply");
  return OFAB_OK;
}
