// Shared device/host helpers for libofab (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ofab.h"

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

// ---- error plumbing (thread-local message, never throws) -------------------------------------
void ofab_set_error(const char* fmt, ...);
int ofab_cuda_fail(cudaError_t e, const char* what);

#define OFAB_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ofab_set_error(__VA_ARGS__);   \
      return OFAB_ERR_ARG;           \
    }                                \
  } while (0)

#define OFAB_LAUNCH_CHECK(what)                      \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return ofab_cuda_fail(e__, what); \
  } while (0)

int ofab_sm_count();  // cached

// tcgen05 attention (attn_tc.cu), dispatched from ofab_attn_fwd / ofab_attn_bwd (attn.cu).  *_fwd / *_bwd return OFAB_OK,
// a negative error, or 1 = "not taken" (the driver refused a tensor map for these strides): run the mma.sync kernels.
int ofab_attn_tc_eligible(const ofab_attn_fwd_args* a);
int ofab_attn_tc_fwd(const ofab_attn_fwd_args* a, ofab_stream_t stream);
int ofab_attn_tc_bwd(const ofab_attn_bwd_args* a, ofab_stream_t stream);

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A step is ~950 short kernels back to back on one stream (CUDA graph): with plain stream order each boundary costs a
// full drain + launch + ramp.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, kernel N+1's CTAs
// become resident as soon as every CTA of kernel N has started (N calls pdl_launch() first thing), run their
// prologue (barrier init, TMEM allocation, descriptor prefetch, index math) and park in pdl_wait() until kernel N has
// COMPLETED and its writes are visible.  Rule for every kernel launched through ofab_launch(): no global-memory read
// or write before pdl_wait().  Ordering stays transitive (a kernel cannot complete before its own wait returns).
bool ofab_pdl_enabled();  // OFAB_PDL=0 in the environment or ofab_set_pdl(0) turns the attribute off
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t ofab_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ofab_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 consecutive elements (16 B for bf16, 32 B for fp32) <-> 8 floats
struct f8 {
  float v[8];
};

__device__ __forceinline__ f8 unpack8(const uint4& u) {  // 8 packed bf16 -> 8 floats
  const bf162* h = reinterpret_cast<const bf162*>(&u);
  f8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ f8 load8(const bf16* p) { return unpack8(*reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ f8 load8(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store8(bf16* p, const f8& r) {
  uint4 u;
  bf162* h = reinterpret_cast<bf162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const f8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// MUFU approximations without the denormal / range fix-up code the CUDA intrinsics carry (3-4 extra
// instructions each): inputs here are O(1), flush-to-zero is harmless.
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// erf-GELU and derivative in fp32 (ofasys/module/gelu.py:18-19 -> F.gelu(x.float())).
// q = 0.5 erfc(|x|/sqrt2) via Abramowitz-Stegun 7.1.26 (|err| <= 0.75e-7 on q, far below the bf16 rounding of
// the result): one rcp + one ex2 + 6 FMA instead of erff's long path; the exponential is shared with the pdf
// term of the gradient.  cdf = 1 - q for x >= 0, q otherwise.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_x) {
  const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  const float e = fast_ex2(x * x * (-0.5f * 1.4426950408889634f));  // exp(-x^2/2)
  const float poly =
      t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f), 0.5f * 1.421413741f), 0.5f * -0.284496736f),
               0.5f * 0.254829592f);
  const float q = poly * e;
  cdf = x >= 0.f ? 1.0f - q : q;
  pdf_x = 0.39894228040143268f * e * x;
}
__device__ __forceinline__ float gelu_f(float x) {
  const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
  const float e = fast_ex2(x * x * (-0.5f * 1.4426950408889634f));
  const float poly =
      t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f), 0.5f * 1.421413741f), 0.5f * -0.284496736f),
               0.5f * 0.254829592f);
  return fmaf(-fabsf(x), poly * e, fmaxf(x, 0.f));  // x * cdf = max(x,0) - |x| q
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, px;
  gelu_parts(x, cdf, px);
  return cdf + px;
}

// ---- packed fp32 (Blackwell FFMA2: two fp32 FMAs per issued instruction) ---------------------------------------------
// The HBM-bound row kernels with a transcendental per element (GELU + LayerNorm) are ISSUE-bound with scalar fp32 math
// (ncu, profiles/r02_ncu_ln_gelu.csv: 69-74 % issue-active at 36 % of HBM peak); pairs of neighbouring columns go through
// fma.rn.f32x2 instead.  Results are bit-identical to the scalar fmaf / * / + on each half (same IEEE operation).
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// q = 0.5 erfc(|x| / sqrt 2) and e = exp(-x^2 / 2) for a pair (same polynomial as gelu_parts)
__device__ __forceinline__ void gelu_qe2(float x0, float x1, float& q0, float& q1, float& e0, float& e1, float& nax0, float& nax1) {
  nax0 = __uint_as_float(__float_as_uint(x0) | 0x80000000u);  // -|x|
  nax1 = __uint_as_float(__float_as_uint(x1) | 0x80000000u);
  float d0, d1, s0, s1, t0, t1, p0, p1;
  const float nk = -0.3275911f * 0.70710678118654752f;
  fma2(d0, d1, nax0, nax1, nk, nk, 1.0f, 1.0f);
  t0 = fast_rcp(d0);
  t1 = fast_rcp(d1);
  mul2(s0, s1, x0, x1, x0, x1);
  const float c = -0.5f * 1.4426950408889634f;
  mul2(s0, s1, s0, s1, c, c);
  e0 = fast_ex2(s0);  // exp(-x^2/2)
  e1 = fast_ex2(s1);
  const float a5 = 0.5f * 1.061405429f, a4 = 0.5f * -1.453152027f, a3 = 0.5f * 1.421413741f, a2 = 0.5f * -0.284496736f,
              a1 = 0.5f * 0.254829592f;
  fma2(p0, p1, t0, t1, a5, a5, a4, a4);
  fma2(p0, p1, t0, t1, p0, p1, a3, a3);
  fma2(p0, p1, t0, t1, p0, p1, a2, a2);
  fma2(p0, p1, t0, t1, p0, p1, a1, a1);
  mul2(p0, p1, t0, t1, p0, p1);
  mul2(q0, q1, p0, p1, e0, e1);
}
// gelu(x) = max(x, 0) - |x| q for a pair of values (value-identical to gelu_f)
__device__ __forceinline__ void gelu2(float& x0, float& x1) {
  float q0, q1, e0, e1, n0, n1;
  gelu_qe2(x0, x1, q0, q1, e0, e1, n0, n1);
  fma2(x0, x1, n0, n1, q0, q1, fmaxf(x0, 0.f), fmaxf(x1, 0.f));
}
// cdf and x * pdf for a pair (value-identical to gelu_parts)
__device__ __forceinline__ void gelu_parts2(float x0, float x1, float& cdf0, float& cdf1, float& px0, float& px1) {
  float q0, q1, e0, e1, n0, n1, m0, m1;
  gelu_qe2(x0, x1, q0, q1, e0, e1, n0, n1);
  fma2(m0, m1, q0, q1, -1.0f, -1.0f, 1.0f, 1.0f);  // 1 - q
  cdf0 = x0 >= 0.f ? m0 : q0;
  cdf1 = x1 >= 0.f ? m1 : q1;
  const float k = 0.39894228040143268f;
  mul2(px0, px1, e0, e1, k, k);
  mul2(px0, px1, px0, px1, x0, x1);
}

// ---- dropout / drop-path masks: counter-based RNG (Philox4x32, 7 rounds), no mask tensor in HBM --------------------
// The mask of the 8-column vector starting at column c of row r is a pure function of (seed, step, site, r, c / 8),
// so forward and backward kernels regenerate the same bits whatever their thread layout.  `state` lives in device
// memory ([0] seed, [1] step counter) so a CUDA-graph replay draws fresh masks: the host only bumps the counter.
// Semantics: ofasys/module/dropout.py:14-25 (F.dropout: zero with probability p, scale kept values by 1/(1-p)) and
// ofasys/module/droppath.py:13-63 (per-sample keep mask scaled by 1/keep on the residual branch).
struct DropArgs {
  const unsigned long long* state;  // device [2]: seed, step
  uint32_t site;                     // call site within the step (host counter, same order every step)
  uint32_t thresh16;                 // an element is dropped when its 16 random bits < thresh16  (= round(p * 65536))
  float inv_keep;                    // 1 / (1 - p)
  uint32_t dp_thresh32;              // drop-path: a sample is dropped when its 32 random bits < dp_thresh32; 0 = off
  float dp_inv_keep;                 // 1 / (1 - drop_path_rate)
  int rows_per_sample;               // rows are [B * T]: sample of row r = r / rows_per_sample
};
// public descriptor -> kernel form.  Returns false (message set) on a bad descriptor.
static inline bool ofab_drop_args(const ofab_dropout* d, DropArgs& a, const char* who) {
  if (d->state == nullptr || !(d->p >= 0.f && d->p < 1.f) || !(d->drop_path >= 0.f && d->drop_path < 1.f) ||
      (d->drop_path > 0.f && d->rows_per_sample <= 0)) {
    ofab_set_error("%s: bad ofab_dropout (state=%p p=%g drop_path=%g rows_per_sample=%d)", who, (const void*)d->state, (double)d->p,
                   (double)d->drop_path, d->rows_per_sample);
    return false;
  }
  a.state = reinterpret_cast<const unsigned long long*>(d->state);
  a.site = d->site;
  a.thresh16 = (uint32_t)(d->p * 65536.0f + 0.5f);
  a.inv_keep = 1.0f / (1.0f - d->p);
  const double t = (double)d->drop_path * 4294967296.0;
  a.dp_thresh32 = d->drop_path > 0.f ? (uint32_t)(t > 4294967295.0 ? 4294967295.0 : (t < 1.0 ? 1.0 : t)) : 0u;
  a.dp_inv_keep = 1.0f / (1.0f - d->drop_path);
  a.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  return true;
}
__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct DropCtx {  // per-kernel constants
  uint2 key;
  uint32_t step_lo;
};
__device__ __forceinline__ DropCtx drop_ctx(const DropArgs& d) {
  const unsigned long long seed = d.state[0], step = d.state[1];
  DropCtx c;
  c.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
  c.step_lo = (uint32_t)step;
  return c;
}
// multipliers (0 or inv_keep [* drop-path scale]) of the 8 columns [c, c+8) of row `row`
__device__ __forceinline__ f8 drop_mask8(const DropArgs& d, const DropCtx& k, int64_t row, int c) {
  float scale = d.inv_keep;
  if (d.dp_thresh32 != 0u) {
    const uint32_t b = (uint32_t)(row / d.rows_per_sample);
    const uint4 r = philox4x32_7(make_uint4(b, 0x0D509A7Fu, d.site, k.step_lo), k.key);
    scale = r.x < d.dp_thresh32 ? 0.f : scale * d.dp_inv_keep;
  }
  const uint4 r = philox4x32_7(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32) ^ (d.site * 0x9E3779B9u), (uint32_t)(c >> 3), k.step_lo), k.key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  f8 m;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t bits = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
    m.v[j] = bits < d.thresh16 ? 0.f : scale;
  }
  return m;
}

// ---- async row pipeline: cp.async.bulk (TMA engine, 1-D) global -> shared ring, mbarrier completion ----------
// HBM-bound row kernels (LayerNorm family) keep several rows per CTA in flight without holding them in registers:
// one elected thread issues a bulk copy per input and stage, all threads wait on the stage's mbarrier and read
// their 16/32 bytes from shared memory.
namespace rowpipe {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void init(uint64_t* bars, int n) {  // call from one thread, then __syncthreads()
  for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bars + i)), "r"(1));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
// bytes % 16 == 0, src and dst 16-byte aligned
__device__ __forceinline__ void load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
               "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = s32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
}  // namespace rowpipe

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
