// libofab core: error plumbing, device checks, small data-movement kernels.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void ofab_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int ofab_cuda_fail(cudaError_t e, const char* what) {
  ofab_set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return OFAB_ERR_CUDA;
}
int ofab_sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int g_pdl = -1;  // -1: read OFAB_PDL from the environment on first use
bool ofab_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("OFAB_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
extern "C" int ofab_set_pdl(int on) {
  const int prev = ofab_pdl_enabled() ? 1 : 0;
  g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int ofab_version(void) { return 100; }
extern "C" const char* ofab_last_error(void) { return g_err; }
extern "C" int ofab_num_sms(void) { return ofab_sm_count(); }
extern "C" int ofab_device_check(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_device_check");
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    ofab_set_error("ofab_device_check: device %d is sm_%d%d; libofab is built for sm_100a only", device, major, minor);
    return OFAB_ERR_DEVICE;
  }
  return OFAB_OK;
}

namespace {
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n8, int64_t n) {
  pdl_launch();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x)
    store8(y + i * 8, load8(x + i * 8));
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (n & ~(int64_t)7) + threadIdx.x;
    y[i] = __float2bfloat16(x[i]);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n8, int64_t n) {
  pdl_launch();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x)
    store8(y + i * 8, load8(x + i * 8));
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t i = (n & ~(int64_t)7) + threadIdx.x;
    y[i] = __bfloat162float(x[i]);
  }
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n4, int64_t n) {
  pdl_launch();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(o)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n & ~(int64_t)3) + threadIdx.x;
    o[i] = a[i] + b[i];
  }
}
__global__ void scale_cols_kernel(const bf16* __restrict__ W, const bf16* __restrict__ c, bf16* __restrict__ out,
                                  int64_t rows, int64_t cols, int group) {
  pdl_launch();
  pdl_wait();
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i % cols;
    out[i] = __float2bfloat16(__bfloat162float(W[i]) * __bfloat162float(c[k / group]));
  }
}
// one block per (row-chunk, head): dW = dWe * c; dc[h] += sum dWe * W
__global__ void scale_cols_bwd_kernel(const bf16* __restrict__ dWe, const bf16* __restrict__ W, const bf16* __restrict__ c,
                                      bf16* __restrict__ dW, float* __restrict__ dc, int64_t rows, int64_t cols, int group) {
  pdl_launch();
  pdl_wait();
  const int h = blockIdx.y;
  const float ch = __bfloat162float(c[h]);
  float s = 0.f;
  const int64_t n = rows * group;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / group, k = (int64_t)h * group + i % group;
    const float g = __bfloat162float(dWe[r * cols + k]);
    s += g * __bfloat162float(W[r * cols + k]);
    dW[r * cols + k] = __float2bfloat16(g * ch);
  }
  s = warp_sum(s);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(dc + h, t);
  }
}
template <typename T>
__global__ void patch_im2col_kernel(const T* __restrict__ img, int B, int C, int H, int W, int p, bf16* __restrict__ cols, int64_t ldk) {
  // one thread per output element; consecutive threads walk pw fastest -> coalesced image reads
  const int gh = H / p, gw = W / p;
  const int64_t kk = (int64_t)C * p * p;
  const int64_t total = (int64_t)B * gh * gw * ldk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ldk;
    const int k = (int)(i % ldk);
    float v = 0.f;
    if (k < kk) {
      const int c = k / (p * p), ph = (k / p) % p, pw = k % p;
      const int b = (int)(row / (gh * gw)), pr = (int)(row % (gh * gw));
      const int y = (pr / gw) * p + ph, x = (pr % gw) * p + pw;
      v = (float)img[(((int64_t)b * C + c) * H + y) * W + x];
    }
    cols[i] = __float2bfloat16(v);
  }
}
__global__ void relu_bwd_inplace_kernel(const bf16* __restrict__ y, bf16* __restrict__ dy, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    f8 a = load8(y + i * 8), g = load8(dy + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] = a.v[j] > 0.f ? g.v[j] : 0.f;
    store8(dy + i * 8, g);
  }
}
__global__ void relu_inplace_kernel(bf16* __restrict__ y, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    f8 a = load8(y + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] = fmaxf(a.v[j], 0.f);
    store8(y + i * 8, a);
  }
}
inline int ew_grid(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace

extern "C" int ofab_cast_f32_bf16(const float* x, void* y, int64_t n, ofab_stream_t stream) {
  if (n <= 0) return OFAB_OK;
  ofab_launch(cast_f32_bf16_kernel, dim3(ew_grid(n / 8 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, x, (bf16*)y, n / 8, n);
  OFAB_LAUNCH_CHECK("ofab_cast_f32_bf16");
  return OFAB_OK;
}
extern "C" int ofab_cast_bf16_f32(const void* x, float* y, int64_t n, ofab_stream_t stream) {
  if (n <= 0) return OFAB_OK;
  ofab_launch(cast_bf16_f32_kernel, dim3(ew_grid(n / 8 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, y, n / 8, n);
  OFAB_LAUNCH_CHECK("ofab_cast_bf16_f32");
  return OFAB_OK;
}
// ---- multi-tensor copy: one launch moves a whole gradient bucket between parameter-shaped tensors and a flat buffer
struct MultiCopyChunk {
  const void* src;
  void* dst;
  unsigned long long bytes;
};
namespace {
__global__ void multi_copy_kernel(const MultiCopyChunk* __restrict__ chunks, int64_t n) {
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const MultiCopyChunk c = chunks[i];
    const char* s = (const char*)c.src;
    char* d = (char*)c.dst;
    const unsigned long long nb = c.bytes;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
      const unsigned long long nv = nb / 16;
      for (unsigned long long j = threadIdx.x; j < nv; j += blockDim.x) ((uint4*)d)[j] = __ldg(((const uint4*)s) + j);
      for (unsigned long long j = nv * 16 + threadIdx.x; j < nb; j += blockDim.x) d[j] = s[j];
    } else {
      for (unsigned long long j = threadIdx.x; j < nb; j += blockDim.x) d[j] = s[j];
    }
  }
}
}  // namespace
extern "C" int ofab_multi_copy(const void* chunks, int64_t n_chunks, ofab_stream_t stream) {
  if (n_chunks <= 0) return OFAB_OK;
  OFAB_REQUIRE(chunks != nullptr, "ofab_multi_copy: null chunk table");
  const int64_t grid = n_chunks < 148 * 16 ? n_chunks : 148 * 16;
  multi_copy_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const MultiCopyChunk*)chunks, n_chunks);
  OFAB_LAUNCH_CHECK("ofab_multi_copy");
  return OFAB_OK;
}
extern "C" int ofab_add_f32(const float* a, const float* b, float* out, int64_t n, ofab_stream_t stream) {
  if (n <= 0) return OFAB_OK;
  ofab_launch(add_f32_kernel, dim3(ew_grid(n / 4 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, a, b, out, n / 4, n);
  OFAB_LAUNCH_CHECK("ofab_add_f32");
  return OFAB_OK;
}
extern "C" int ofab_scale_cols(const void* W, const void* c, void* out, int64_t rows, int64_t cols, int group, ofab_stream_t stream) {
  OFAB_REQUIRE(group > 0 && cols % group == 0, "ofab_scale_cols: cols %% group != 0");
  ofab_launch(scale_cols_kernel, dim3(ew_grid(rows * cols, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)W, (const bf16*)c, (bf16*)out, rows, cols, group);
  OFAB_LAUNCH_CHECK("ofab_scale_cols");
  return OFAB_OK;
}
extern "C" int ofab_scale_cols_bwd(const void* dW_eff, const void* W, const void* c, void* dW, float* dc, int64_t rows,
                                   int64_t cols, int group, ofab_stream_t stream) {
  OFAB_REQUIRE(group > 0 && cols % group == 0, "ofab_scale_cols_bwd: cols %% group != 0");
  dim3 grid((unsigned)ew_grid(rows * group, 256 * 8), (unsigned)(cols / group));
  ofab_launch(scale_cols_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dW_eff, (const bf16*)W, (const bf16*)c, (bf16*)dW, dc, rows, cols, group);
  OFAB_LAUNCH_CHECK("ofab_scale_cols_bwd");
  return OFAB_OK;
}
extern "C" int ofab_patch_im2col(const void* img, int img_dt, int B, int C, int H, int W, int p, void* cols, int64_t ldk,
                                 ofab_stream_t stream) {
  OFAB_REQUIRE(p > 0 && H % p == 0 && W % p == 0, "ofab_patch_im2col: image %dx%d not divisible by patch %d", H, W, p);
  OFAB_REQUIRE(ldk >= (int64_t)C * p * p && ldk % 8 == 0, "ofab_patch_im2col: ldk=%lld must be >= C*p*p and a multiple of 8", (long long)ldk);
  const int64_t total = (int64_t)B * (H / p) * (W / p) * ldk;
  if (img_dt == OFAB_F32)
    patch_im2col_kernel<float><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)img, B, C, H, W, p, (bf16*)cols, ldk);
  else
    patch_im2col_kernel<bf16><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)img, B, C, H, W, p, (bf16*)cols, ldk);
  OFAB_LAUNCH_CHECK("ofab_patch_im2col");
  return OFAB_OK;
}
extern "C" int ofab_relu_bwd_inplace(const void* y, void* dy, int64_t n, ofab_stream_t stream) {
  OFAB_REQUIRE(n % 8 == 0, "ofab_relu_bwd_inplace: n %% 8 != 0");
  relu_bwd_inplace_kernel<<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (bf16*)dy, n / 8);
  OFAB_LAUNCH_CHECK("ofab_relu_bwd_inplace");
  return OFAB_OK;
}
extern "C" int ofab_relu_inplace(void* y, int64_t n, ofab_stream_t stream) {
  OFAB_REQUIRE(n % 8 == 0, "ofab_relu_inplace: n %% 8 != 0");
  relu_inplace_kernel<<<ew_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((bf16*)y, n / 8);
  OFAB_LAUNCH_CHECK("ofab_relu_inplace");
  return OFAB_OK;
}

// ---- standalone dropout (un-fused call sites: the reference-layout API path, tests) ------------------------------
namespace {
template <typename T>
__global__ void __launch_bounds__(256) dropout_apply_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int cols, const DropArgs da) {
  pdl_launch();
  pdl_wait();
  const DropCtx dk = drop_ctx(da);
  const int vpr = cols >> 3;  // 8-column vectors per row
  const int64_t nvec = rows * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vpr;
    const int c = (int)(i % vpr) * 8;
    f8 v = load8(x + row * cols + c);
    const f8 m = drop_mask8(da, dk, row, c);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] *= m.v[j];
    store8(y + row * cols + c, v);
  }
}
}  // namespace

extern "C" int ofab_dropout_apply(const void* x, void* y, int dt, int64_t rows, int cols, const ofab_dropout* drop, ofab_stream_t stream) {
  OFAB_REQUIRE(x != nullptr && y != nullptr && drop != nullptr, "ofab_dropout_apply: NULL argument");
  OFAB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, "ofab_dropout_apply: cols=%d must be a positive multiple of 8", cols);
  OFAB_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)y) & 15) == 0, "ofab_dropout_apply: x / y must be 16-byte aligned");
  DropArgs da{};
  if (!ofab_drop_args(drop, da, "ofab_dropout_apply")) return OFAB_ERR_ARG;
  if (rows == 0) return OFAB_OK;
  const int64_t nvec = rows * (cols >> 3);
  const int64_t want = (nvec + 255) / 256, cap = (int64_t)ofab_sm_count() * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (dt == OFAB_F32)
    ofab_launch((dropout_apply_kernel<float>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)x, (float*)y, rows, cols, da);
  else if (dt == OFAB_BF16)
    ofab_launch((dropout_apply_kernel<bf16>), dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, rows, cols, da);
  else {
    ofab_set_error("ofab_dropout_apply: dt=%d", dt);
    return OFAB_ERR_ARG;
  }
  OFAB_LAUNCH_CHECK("ofab_dropout_apply");
  return OFAB_OK;
}
