// Audio fbank subsampler pieces (ofasys/module/subsample.py:27-63) around the tcgen05 GEMM:
//   conv1 (1 -> C, 3x3, stride 2) + ReLU : direct CUDA-core kernel, channel-last output
//   conv2 (C -> C, 3x3, stride 2)        : im2col (vector copies) -> gemm.cu -> ReLU
// plus the [O, A, B] -> [O, B, A] weight permute that maps the reference's NCHW weight / feature
// order onto the channel-last order used here.
#include "common.cuh"

namespace {

template <typename TIN>
__global__ void __launch_bounds__(128) conv1_relu_fwd_kernel(const TIN* __restrict__ x, int B, int L, int F, const bf16* __restrict__ w,
                                                             const bf16* __restrict__ bias, int C, bf16* __restrict__ out, int H1, int W1) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  extern __shared__ float xs[];  // 3 x F
  const int b = blockIdx.x / H1, h = blockIdx.x % H1;
  for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) xs[i] = (float)x[((int64_t)b * L + 2 * h + i / F) * F + i % F];
  __syncthreads();
  for (int cg = threadIdx.x; cg * 8 < C; cg += blockDim.x) {
    float wr[8][9], br[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      br[j] = __bfloat162float(bias[cg * 8 + j]);
#pragma unroll
      for (int k = 0; k < 9; ++k) wr[j][k] = __bfloat162float(w[(cg * 8 + j) * 9 + k]);
    }
    for (int wo = 0; wo < W1; ++wo) {
      float in[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) in[k] = xs[(k / 3) * F + 2 * wo + k % 3];
      f8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = br[j];
#pragma unroll
        for (int k = 0; k < 9; ++k) a += wr[j][k] * in[k];
        o.v[j] = fmaxf(a, 0.f);
      }
      store8(out + (((int64_t)b * H1 + h) * W1 + wo) * C + cg * 8, o);
    }
  }
}

// persistent: each block walks (b, h) rows, accumulating dw/db of its threads' channels in registers
template <typename TIN>
__global__ void __launch_bounds__(128) conv1_relu_bwd_kernel(const TIN* __restrict__ x, int B, int L, int F, const bf16* __restrict__ y,
                                                             const bf16* __restrict__ dy, int C, float* __restrict__ dw, float* __restrict__ db,
                                                             int H1, int W1) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  extern __shared__ float xs[];
  const int cg = threadIdx.x;  // one 8-channel group per thread; host guarantees C/8 <= blockDim.x
  const bool live = cg * 8 < C;
  float aw[8][9], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ab[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) aw[j][k] = 0.f;
  }
  for (int rowi = blockIdx.x; rowi < B * H1; rowi += gridDim.x) {
    const int b = rowi / H1, h = rowi % H1;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) xs[i] = (float)x[((int64_t)b * L + 2 * h + i / F) * F + i % F];
    __syncthreads();
    if (live) {
      for (int wo = 0; wo < W1; ++wo) {
        const int64_t off = (((int64_t)b * H1 + h) * W1 + wo) * C + cg * 8;
        const f8 yy = load8(y + off), gg = load8(dy + off);
        float in[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) in[k] = xs[(k / 3) * F + 2 * wo + k % 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = yy.v[j] > 0.f ? gg.v[j] : 0.f;
          ab[j] += g;
#pragma unroll
          for (int k = 0; k < 9; ++k) aw[j][k] += g * in[k];
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(db + cg * 8 + j, ab[j]);
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(dw + (cg * 8 + j) * 9 + k, aw[j][k]);
    }
  }
}

// cols[(b, ho, wo), (kh*3+kw)*C + c] = x[b, 2ho+kh, 2wo+kw, c]
__global__ void im2col_3x3s2_kernel(const bf16* __restrict__ x, int B, int Hin, int Win, int C, bf16* __restrict__ cols, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * 9 * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int k = (int)((i / C8) % 9);
    const int64_t r = i / ((int64_t)C8 * 9);
    const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((int64_t)Wo * Ho));
    const uint4 v = *reinterpret_cast<const uint4*>(x + (((int64_t)b * Hin + 2 * ho + k / 3) * Win + 2 * wo + k % 3) * C + c8 * 8);
    *reinterpret_cast<uint4*>(cols + (r * 9 + k) * C + c8 * 8) = v;
  }
}
// dx[b, h, w, c] = sum over (kh, kw) with h = 2ho+kh, w = 2wo+kw of dcols[(b,ho,wo), (kh*3+kw)*C + c]
__global__ void col2im_3x3s2_kernel(const bf16* __restrict__ dcols, int B, int Hin, int Win, int C, bf16* __restrict__ dx, int Ho, int Wo) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  const int C8 = C / 8;
  const int64_t total = (int64_t)B * Hin * Win * C8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const int64_t r = i / C8;
    const int w = (int)(r % Win), h = (int)((r / Win) % Hin), b = (int)(r / ((int64_t)Win * Hin));
    f8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = h - kh;
      if (hh < 0 || (hh & 1) || hh / 2 >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = w - kw;
        if (ww < 0 || (ww & 1) || ww / 2 >= Wo) continue;
        const f8 v = load8(dcols + ((((int64_t)b * Ho + hh / 2) * Wo + ww / 2) * 9 + kh * 3 + kw) * C + c8 * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j];
      }
    }
    store8(dx + r * C + c8 * 8, acc);
  }
}
// out[o, b, a] = in[o, a, b]   (32x32 smem tiles)
__global__ void transpose_last2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int A, int Bd) {
  pdl_launch();
  pdl_wait();  // PDL: no global access above
  __shared__ bf16 tile[32][33];
  const int64_t o = blockIdx.z;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int a = a0 + i, b = b0 + threadIdx.x;
    if (a < A && b < Bd) tile[i][threadIdx.x] = in[(o * A + a) * Bd + b];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int b = b0 + i, a = a0 + threadIdx.x;
    if (a < A && b < Bd) out[(o * Bd + b) * A + a] = tile[threadIdx.x][i];
  }
}
inline int ew_grid(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace

extern "C" int ofab_conv1_relu_fwd(const void* fbank, int in_dt, int B, int L, int F, const void* w, const void* bias, int C, void* out,
                                   ofab_stream_t stream) {
  OFAB_REQUIRE(L >= 3 && F >= 3 && C % 8 == 0, "ofab_conv1_relu_fwd: bad shape L=%d F=%d C=%d", L, F, C);
  const int H1 = (L - 3) / 2 + 1, W1 = (F - 3) / 2 + 1;
  const int smem = 3 * F * 4;
  if (in_dt == OFAB_F32)
    ofab_launch(conv1_relu_fwd_kernel<float>, dim3(B * H1), dim3(128), (size_t)(smem), (cudaStream_t)stream, (const float*)fbank, B, L, F, (const bf16*)w, (const bf16*)bias, C, (bf16*)out, H1, W1);
  else
    ofab_launch(conv1_relu_fwd_kernel<bf16>, dim3(B * H1), dim3(128), (size_t)(smem), (cudaStream_t)stream, (const bf16*)fbank, B, L, F, (const bf16*)w, (const bf16*)bias, C, (bf16*)out, H1, W1);
  OFAB_LAUNCH_CHECK("ofab_conv1_relu_fwd");
  return OFAB_OK;
}
extern "C" int ofab_conv1_relu_bwd(const void* fbank, int in_dt, int B, int L, int F, const void* y, const void* dy, int C, float* dw,
                                   float* db, ofab_stream_t stream) {
  OFAB_REQUIRE(L >= 3 && F >= 3 && C % 8 == 0 && C <= 1024, "ofab_conv1_relu_bwd: bad shape L=%d F=%d C=%d (C <= 1024)", L, F, C);
  const int H1 = (L - 3) / 2 + 1, W1 = (F - 3) / 2 + 1;
  const int smem = 3 * F * 4;
  const int grid = B * H1 < 592 ? B * H1 : 592;
  if (in_dt == OFAB_F32)
    ofab_launch(conv1_relu_bwd_kernel<float>, dim3(grid), dim3(128), (size_t)(smem), (cudaStream_t)stream, (const float*)fbank, B, L, F, (const bf16*)y, (const bf16*)dy, C, dw, db, H1, W1);
  else
    ofab_launch(conv1_relu_bwd_kernel<bf16>, dim3(grid), dim3(128), (size_t)(smem), (cudaStream_t)stream, (const bf16*)fbank, B, L, F, (const bf16*)y, (const bf16*)dy, C, dw, db, H1, W1);
  OFAB_LAUNCH_CHECK("ofab_conv1_relu_bwd");
  return OFAB_OK;
}
extern "C" int ofab_im2col_3x3s2(const void* x, int B, int Hin, int Win, int C, void* cols, ofab_stream_t stream) {
  OFAB_REQUIRE(Hin >= 3 && Win >= 3 && C % 8 == 0, "ofab_im2col_3x3s2: bad shape");
  const int Ho = (Hin - 3) / 2 + 1, Wo = (Win - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * 9 * (C / 8);
  ofab_launch(im2col_3x3s2_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)x, B, Hin, Win, C, (bf16*)cols, Ho, Wo);
  OFAB_LAUNCH_CHECK("ofab_im2col_3x3s2");
  return OFAB_OK;
}
extern "C" int ofab_col2im_3x3s2(const void* dcols, int B, int Hin, int Win, int C, void* dx, ofab_stream_t stream) {
  OFAB_REQUIRE(Hin >= 3 && Win >= 3 && C % 8 == 0, "ofab_col2im_3x3s2: bad shape");
  const int Ho = (Hin - 3) / 2 + 1, Wo = (Win - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Hin * Win * (C / 8);
  ofab_launch(col2im_3x3s2_kernel, dim3(ew_grid(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const bf16*)dcols, B, Hin, Win, C, (bf16*)dx, Ho, Wo);
  OFAB_LAUNCH_CHECK("ofab_col2im_3x3s2");
  return OFAB_OK;
}
extern "C" int ofab_transpose_last2(const void* in, void* out, int64_t O, int A, int Bd, ofab_stream_t stream) {
  OFAB_REQUIRE(O > 0 && O < 65536 && A > 0 && Bd > 0, "ofab_transpose_last2: bad shape O=%lld A=%d B=%d (O < 65536)", (long long)O, A, Bd);
  dim3 grid((Bd + 31) / 32, (A + 31) / 32, (unsigned)O), block(32, 8);
  ofab_launch(transpose_last2_kernel, dim3(grid), dim3(block), (size_t)(0), (cudaStream_t)stream, (const bf16*)in, (bf16*)out, A, Bd);
  OFAB_LAUNCH_CHECK("ofab_transpose_last2");
  return OFAB_OK;
}
