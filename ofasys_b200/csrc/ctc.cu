// CTC criterion of the ASR recipe (second half): the loss the reference adds to the cross entropy for
// speech (ofasys/engine/criterion/speech_to_text_loss.py:206-237,339-379): logits = F.linear(encoder_out, E[phone range]),
// lprobs = log_softmax(logits.float()), F.ctc_loss(lprobs, targets, input_lengths, target_lengths, blank, reduction="sum",
// zero_infinity).  The algorithm is PyTorch's (third-party; torch 2.11 aten/src/ATen/native/LossCTC.cpp: the alpha / beta
// recursions of Graves et al. 2006 in log space, eq. 16 for the gradient).  Here the log-softmax is fused: the kernels
// read the raw logits, keep log-sum-exp per frame, and emit the gradient with respect to the LOGITS directly
//   d nll / d x[t, c] = softmax[t, c] - exp(logsum_{s: l'_s = c}(alpha_t(s) + beta_t(s)) + nll - lprob[t, c])
// One CTA per utterance (the time recursion is sequential; T' ~ 250 frames, S = 2 L + 1 <= 1025 states): the problem is
// tiny next to the model step, the point is that no [T, B, C] fp32 log-prob tensor and no host round trip exist.
#include "common.cuh"

namespace {

constexpr int kCtcThreads = 256;
constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float logadd(float a, float b) {
  if (a == kNegInf) return b;
  if (b == kNegInf) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float block_max256(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < kCtcThreads / 32; ++w) t = fmaxf(t, red[w]);
  return t;
}
__device__ __forceinline__ float block_sum256c(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kCtcThreads / 32; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ void atomic_logadd(float* addr, float v) {  // shared-memory log-add-exp (repeated labels only collide)
  if (v == kNegInf) return;
  int* ia = reinterpret_cast<int*>(addr);
  int old = *ia, assumed;
  do {
    assumed = old;
    old = atomicCAS(ia, assumed, __float_as_int(logadd(__int_as_float(assumed), v)));
  } while (old != assumed);
}

template <typename T>
struct CtcIn {
  const T* x;            // logits, element (t, b, c) at x[t * ts + b * bs + c]
  int64_t ts, bs;
  int Tmax, C;
  const int64_t* in_len;   // [B] frames per utterance (NULL: Tmax)
  const int64_t* targets;  // [B, Lmax] labels, left-aligned
  const int64_t* tgt_len;  // [B]
  int Lmax, blank;
};

template <typename T>
__device__ __forceinline__ int ext_label(const CtcIn<T>& a, int b, int s) {  // l'_s: blank at even s, label (s-1)/2 at odd s
  return (s & 1) ? (int)a.targets[(int64_t)b * a.Lmax + (s >> 1)] : a.blank;
}

// forward: lse[b, t], alpha[b, t, s], nll[b]
template <typename T>
__global__ void __launch_bounds__(kCtcThreads) ctc_alpha_kernel(const CtcIn<T> a, float* __restrict__ lse, float* __restrict__ alpha, int Smax,
                                                                float* __restrict__ nll) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int Tb = a.in_len != nullptr ? (int)min((int64_t)a.Tmax, a.in_len[b]) : a.Tmax;
  const int L = (int)a.tgt_len[b], S = 2 * L + 1;
  float* al = alpha + (int64_t)b * a.Tmax * Smax;
  for (int t = 0; t < Tb; ++t) {  // log-sum-exp of every frame (the fused log_softmax)
    const T* row = a.x + (int64_t)t * a.ts + (int64_t)b * a.bs;
    float m = kNegInf;
    for (int c = tid; c < a.C; c += kCtcThreads) m = fmaxf(m, (float)row[c]);
    m = block_max256(m, red);
    float s = 0.f;
    for (int c = tid; c < a.C; c += kCtcThreads) s += expf((float)row[c] - m);
    s = block_sum256c(s, red);
    if (tid == 0) lse[(int64_t)b * a.Tmax + t] = m + logf(s);
  }
  __syncthreads();
  if (Tb == 0) {
    if (tid == 0) nll[b] = L == 0 ? 0.f : INFINITY;
    return;
  }
  for (int s = tid; s < S; s += kCtcThreads) {
    float v = kNegInf;
    if (s < 2) v = (float)a.x[(int64_t)b * a.bs + ext_label(a, b, s)] - lse[(int64_t)b * a.Tmax];
    al[s] = v;
  }
  for (int t = 1; t < Tb; ++t) {
    __syncthreads();  // alpha[t-1] complete (global memory, same CTA)
    const T* row = a.x + (int64_t)t * a.ts + (int64_t)b * a.bs;
    const float l = lse[(int64_t)b * a.Tmax + t];
    const float* prev = al + (int64_t)(t - 1) * Smax;
    float* cur = al + (int64_t)t * Smax;
    for (int s = tid; s < S; s += kCtcThreads) {
      const int c = ext_label(a, b, s);
      float v = prev[s];
      if (s >= 1) v = logadd(v, prev[s - 1]);
      if (s >= 2 && c != a.blank && c != ext_label(a, b, s - 2)) v = logadd(v, prev[s - 2]);
      cur[s] = v == kNegInf ? kNegInf : v + ((float)row[c] - l);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float* last = al + (int64_t)(Tb - 1) * Smax;
    float v = last[S - 1];
    if (S >= 2) v = logadd(v, last[S - 2]);
    nll[b] = -v;
  }
}

// backward: beta recursion + gradient with respect to the logits (scaled by gscale[0])
template <typename T>
__global__ void __launch_bounds__(kCtcThreads) ctc_beta_grad_kernel(const CtcIn<T> a, const float* __restrict__ lse, const float* __restrict__ alpha,
                                                                    int Smax, const float* __restrict__ nll, const float* __restrict__ gscale,
                                                                    int zero_infinity, T* __restrict__ grad /* same strides as x */) {
  extern __shared__ float sm[];  // beta[2][Smax] | acc[C]
  float* beta0 = sm;
  float* beta1 = sm + Smax;
  float* acc = sm + 2 * Smax;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int Tb = a.in_len != nullptr ? (int)min((int64_t)a.Tmax, a.in_len[b]) : a.Tmax;
  const int L = (int)a.tgt_len[b], S = 2 * L + 1;
  const float nl = nll[b];
  const bool dead = zero_infinity && (nl == INFINITY);
  const float g = dead ? 0.f : gscale[0];
  const float* al = alpha + (int64_t)b * a.Tmax * Smax;
  for (int t = a.Tmax - 1; t >= Tb; --t) {  // frames past the utterance: zero gradient
    T* grow = grad + (int64_t)t * a.ts + (int64_t)b * a.bs;
    for (int c = tid; c < a.C; c += kCtcThreads) grow[c] = (T)0.f;
  }
  for (int t = Tb - 1; t >= 0; --t) {
    const T* row = a.x + (int64_t)t * a.ts + (int64_t)b * a.bs;
    T* grow = grad + (int64_t)t * a.ts + (int64_t)b * a.bs;
    const float l = lse[(int64_t)b * a.Tmax + t];
    float* cur = (t & 1) ? beta1 : beta0;
    const float* nxt = (t & 1) ? beta0 : beta1;
    for (int c = tid; c < a.C; c += kCtcThreads) acc[c] = kNegInf;
    __syncthreads();  // acc cleared; beta[t+1] complete
    for (int s = tid; s < S; s += kCtcThreads) {
      const int c = ext_label(a, b, s);
      const float lp = (float)row[c] - l;
      float v;
      if (t == Tb - 1) {
        v = (s == S - 1 || s == S - 2) ? lp : kNegInf;
      } else {
        v = nxt[s];
        if (s + 1 < S) v = logadd(v, nxt[s + 1]);
        if (s + 2 < S && c != a.blank && c != ext_label(a, b, s + 2)) v = logadd(v, nxt[s + 2]);
        v = v == kNegInf ? kNegInf : v + lp;
      }
      cur[s] = v;
      const float ab = al[(int64_t)t * Smax + s];
      if (ab != kNegInf && v != kNegInf) atomic_logadd(acc + c, ab + v);
    }
    __syncthreads();
    for (int c = tid; c < a.C; c += kCtcThreads) {
      const float lp = (float)row[c] - l;
      float r = expf(lp);
      if (acc[c] != kNegInf) r -= expf(acc[c] + nl - lp);  // alpha * beta carries lprob twice
      grow[c] = (T)(dead ? 0.f : r * g);
    }
    __syncthreads();
  }
}

template <typename T>
int ctc_run(const ofab_ctc_args* p, bool backward, cudaStream_t st) {
  CtcIn<T> a;
  a.x = (const T*)p->logits; a.ts = p->t_stride; a.bs = p->b_stride; a.Tmax = p->T; a.C = p->C;
  a.in_len = p->input_lengths; a.targets = p->targets; a.tgt_len = p->target_lengths; a.Lmax = p->Lmax; a.blank = p->blank;
  const int Smax = 2 * p->Lmax + 1;
  if (!backward) {
    ctc_alpha_kernel<T><<<p->B, kCtcThreads, 0, st>>>(a, p->lse, p->alpha, Smax, p->nll);
  } else {
    const size_t smem = (size_t)(2 * Smax + p->C) * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ctc_beta_grad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return ofab_cuda_fail(e, "ofab_ctc_bwd: shared memory");
    }
    ctc_beta_grad_kernel<T><<<p->B, kCtcThreads, smem, st>>>(a, p->lse, p->alpha, Smax, p->nll, p->gscale, p->zero_infinity, (T*)p->dlogits);
  }
  return OFAB_OK;
}

int ctc_check(const ofab_ctc_args* p, const char* who) {
  OFAB_REQUIRE(p != nullptr && p->logits && p->targets && p->target_lengths && p->lse && p->alpha && p->nll, "%s: NULL argument", who);
  OFAB_REQUIRE(p->B > 0 && p->T > 0 && p->C > 1 && p->Lmax >= 0 && p->blank >= 0 && p->blank < p->C, "%s: bad shape B=%d T=%d C=%d Lmax=%d blank=%d", who, p->B, p->T, p->C, p->Lmax, p->blank);
  OFAB_REQUIRE((size_t)(2 * (2 * p->Lmax + 1) + p->C) * sizeof(float) <= 200 * 1024, "%s: 2 * (2 Lmax + 1) + C floats must fit in shared memory", who);
  OFAB_REQUIRE(p->dt == OFAB_F32 || p->dt == OFAB_BF16, "%s: dt", who);
  return OFAB_OK;
}

}  // namespace

extern "C" int ofab_ctc_fwd(const ofab_ctc_args* p, ofab_stream_t stream) {
  int rc = ctc_check(p, "ofab_ctc_fwd");
  if (rc) return rc;
  rc = p->dt == OFAB_F32 ? ctc_run<float>(p, false, (cudaStream_t)stream) : ctc_run<bf16>(p, false, (cudaStream_t)stream);
  if (rc) return rc;
  OFAB_LAUNCH_CHECK("ofab_ctc_fwd");
  return OFAB_OK;
}

extern "C" int ofab_ctc_bwd(const ofab_ctc_args* p, ofab_stream_t stream) {
  int rc = ctc_check(p, "ofab_ctc_bwd");
  if (rc) return rc;
  OFAB_REQUIRE(p->dlogits != nullptr && p->gscale != nullptr, "ofab_ctc_bwd: dlogits / gscale NULL");
  rc = p->dt == OFAB_F32 ? ctc_run<float>(p, true, (cudaStream_t)stream) : ctc_run<bf16>(p, true, (cudaStream_t)stream);
  if (rc) return rc;
  OFAB_LAUNCH_CHECK("ofab_ctc_bwd");
  return OFAB_OK;
}
