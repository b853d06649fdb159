// Optimizer step that follows the measured path every update: gradient norm, clipping, and the
// fp32-master Adam update of ALL parameters in two multi-tensor launches (HBM-bound: 2 B gradient + 12 B state read,
// 12 B state + 2 B parameter written per element).
//
// Replaces, for bf16 training, the chain  fp16_optimizer.py:104-204 (bf16 grads -> fp32 copies, deferred multiply
// factor, clip coefficient, fp32 step, copy back to bf16)  +  module/utils.py:342-384 (clip_grad_norm_)  +
// engine/optim/adam.py:150-216 (Adam with decoupled weight decay): ~10 elementwise torch kernels per parameter tensor
// there (apex's fused_adam / multi_tensor_l2norm are optional and absent), one pass over the state here.  The clip
// coefficient is computed on the device from the norm, so the step needs no host synchronisation.
#include "common.cuh"

namespace {

constexpr int kChunk = 16384;  // elements of one tensor handled by one CTA
constexpr int kThreads = 256;

struct AdamTensor {  // == ofab_adam_tensor
  bf16* p;
  const bf16* g;  // NULL: unused parameter, zero gradient (fp16_optimizer.py:129-130)
  float* master;
  float* m;
  float* v;
  int64_t n;
  int64_t first_block;  // index of this tensor's first CTA; tensors are sorted by it
};

// which tensor does CTA `blk` work on: last t with first_block <= blk
__device__ __forceinline__ int find_tensor(const AdamTensor* __restrict__ ts, int n_tensors, int64_t blk) {
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ts[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float block_sum256(float v, float* sm /* [8] */) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) t += sm[w];
  return t;
}

__global__ void __launch_bounds__(kThreads) adam_sqnorm_kernel(const AdamTensor* __restrict__ ts, int n_tensors, float* __restrict__ partial) {
  __shared__ float sm[8];
  const AdamTensor t = ts[find_tensor(ts, n_tensors, blockIdx.x)];
  const int64_t off = ((int64_t)blockIdx.x - t.first_block) * kChunk;
  const int64_t n = min((int64_t)kChunk, t.n - off);
  float s = 0.f;
  if (t.g != nullptr) {
    const bf16* g = t.g + off;
    if ((((uintptr_t)g) & 15) == 0) {
      const int64_t n8 = n >> 3;
      for (int64_t i = threadIdx.x; i < n8; i += kThreads) {
        const f8 x = load8(g + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(x.v[j], x.v[j], s);
      }
      for (int64_t i = (n8 << 3) + threadIdx.x; i < n; i += kThreads) {
        const float x = __bfloat162float(g[i]);
        s = fmaf(x, x, s);
      }
    } else {
      for (int64_t i = threadIdx.x; i < n; i += kThreads) {
        const float x = __bfloat162float(g[i]);
        s = fmaf(x, x, s);
      }
    }
  }
  s = block_sum256(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// norm[0] = sqrt(sum of the per-CTA partial sums), accumulated in double (deterministic: fixed order)
__global__ void __launch_bounds__(1024) adam_norm_final_kernel(const float* __restrict__ partial, int64_t n, float* __restrict__ norm) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t += sm[w];
    norm[0] = (float)sqrt(t);
  }
}

struct AdamHyper {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps;
  float step_size;  // lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (adam.py:205-207, evaluated in double on the host)
  float wd_lr;      // weight_decay * lr                         (adam.py:209-210)
  float grad_scale; // the deferred multiply_grads factor        (fp16_optimizer.py:170-172)
  float max_norm;   // <= 0: no clipping
  const float* norm;  // device: unscaled gradient norm (adam_norm_final_kernel) or NULL when max_norm <= 0
};

__device__ __forceinline__ void adam_elem(float g, float& p, float& m, float& v, const AdamHyper& h, float factor) {
  g *= factor;                                        // fp16_optimizer.py:152-168: ONE multiply by c * clip_coef
  m = fmaf(h.one_minus_beta1, g, m * h.beta1);        // adam.py:195
  v = fmaf(h.one_minus_beta2 * g, g, v * h.beta2);    // adam.py:196
  const float denom = sqrtf(v) + h.eps;               // adam.py:203
  if (h.wd_lr != 0.f) p = fmaf(-h.wd_lr, p, p);       // adam.py:209-210
  p = fmaf(-h.step_size, m / denom, p);               // adam.py:212
}

__global__ void __launch_bounds__(kThreads) adam_step_kernel(const AdamTensor* __restrict__ ts, int n_tensors, const AdamHyper h) {
  const AdamTensor t = ts[find_tensor(ts, n_tensors, blockIdx.x)];
  const int64_t off = ((int64_t)blockIdx.x - t.first_block) * kChunk;
  const int64_t n = min((int64_t)kChunk, t.n - off);
  float factor = h.grad_scale;
  if (h.max_norm > 0.f) {
    const float grad_norm = h.grad_scale * h.norm[0];                      // fp16_optimizer.py:178
    factor *= fminf(h.max_norm / (grad_norm + 1e-6f), 1.0f);               // :185-187
  }
  bf16* p = t.p + off;
  const bf16* g = t.g != nullptr ? t.g + off : nullptr;
  float *ma = t.master + off, *m = t.m + off, *v = t.v + off;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g)) & 15) == 0 && ((((uintptr_t)ma) | ((uintptr_t)m) | ((uintptr_t)v)) & 31) == 0;
  const int64_t n8 = aligned ? (n >> 3) : 0;
  for (int64_t i = threadIdx.x; i < n8; i += kThreads) {
    f8 gg, pp = load8(ma + i * 8), mm = load8(m + i * 8), vv = load8(v + i * 8);
    if (g != nullptr) gg = load8(g + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) adam_elem(g != nullptr ? gg.v[j] : 0.f, pp.v[j], mm.v[j], vv.v[j], h, factor);
    store8(ma + i * 8, pp);
    store8(m + i * 8, mm);
    store8(v + i * 8, vv);
    store8(p + i * 8, pp);  // masters -> bf16 parameters (fp16_optimizer.py:134-150)
  }
  for (int64_t i = (n8 << 3) + threadIdx.x; i < n; i += kThreads) {
    float pp = ma[i], mm = m[i], vv = v[i];
    adam_elem(g != nullptr ? __bfloat162float(g[i]) : 0.f, pp, mm, vv, h, factor);
    ma[i] = pp;
    m[i] = mm;
    v[i] = vv;
    p[i] = __float2bfloat16(pp);
  }
}

}  // namespace

extern "C" int ofab_adam_chunk_elems(void) { return kChunk; }

extern "C" int ofab_grad_norm(const ofab_adam_tensor* tensors, int n_tensors, int64_t n_blocks, float* partial, float* norm,
                              ofab_stream_t stream) {
  static_assert(sizeof(ofab_adam_tensor) == sizeof(AdamTensor), "ofab_adam_tensor layout");
  OFAB_REQUIRE(tensors != nullptr && n_tensors > 0 && n_blocks > 0 && n_blocks < (1ll << 31) && partial != nullptr && norm != nullptr,
               "ofab_grad_norm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  adam_sqnorm_kernel<<<(unsigned)n_blocks, kThreads, 0, st>>>(reinterpret_cast<const AdamTensor*>(tensors), n_tensors, partial);
  OFAB_LAUNCH_CHECK("ofab_grad_norm");
  adam_norm_final_kernel<<<1, 1024, 0, st>>>(partial, n_blocks, norm);
  OFAB_LAUNCH_CHECK("ofab_grad_norm final");
  return OFAB_OK;
}

extern "C" int ofab_adam_step(const ofab_adam_tensor* tensors, int n_tensors, int64_t n_blocks, const ofab_adam_hyper* hy,
                              ofab_stream_t stream) {
  OFAB_REQUIRE(tensors != nullptr && n_tensors > 0 && n_blocks > 0 && n_blocks < (1ll << 31) && hy != nullptr, "ofab_adam_step: bad arguments");
  OFAB_REQUIRE(hy->max_norm <= 0.f || hy->norm != nullptr, "ofab_adam_step: max_norm > 0 needs the norm from ofab_grad_norm");
  OFAB_REQUIRE(hy->step >= 1, "ofab_adam_step: step is the 1-based update count");
  AdamHyper h;
  h.beta1 = hy->beta1;
  h.beta2 = hy->beta2;
  h.one_minus_beta1 = (float)(1.0 - (double)hy->beta1_d);
  h.one_minus_beta2 = (float)(1.0 - (double)hy->beta2_d);
  h.eps = hy->eps;
  const double bc1 = 1.0 - pow(hy->beta1_d, (double)hy->step), bc2 = 1.0 - pow(hy->beta2_d, (double)hy->step);
  h.step_size = (float)(hy->lr * sqrt(bc2) / bc1);
  h.wd_lr = (float)(hy->weight_decay * hy->lr);
  h.grad_scale = hy->grad_scale;
  h.max_norm = hy->max_norm;
  h.norm = hy->norm;
  adam_step_kernel<<<(unsigned)n_blocks, kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamTensor*>(tensors), n_tensors, h);
  OFAB_LAUNCH_CHECK("ofab_adam_step");
  return OFAB_OK;
}
