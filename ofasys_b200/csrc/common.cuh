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

__device__ __forceinline__ f8 load8(const bf16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
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

// erf-GELU and derivative in fp32 (ofasys/module/gelu.py:18-19 -> F.gelu(x.float())).
// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below the bf16 rounding of the result): one
// exp + 5 FMA + 1 rcp instead of erff's long path; the exp is shared with the pdf term of the gradient.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));  // MUFU.RCP
  const float e = __expf(-z * z);  // = exp(-x^2/2)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float erf_abs = fmaf(-poly, e, 1.0f);
  cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  pdf_x = 0.39894228040143268f * e * x;
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, px;
  gelu_parts(x, cdf, px);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, px;
  gelu_parts(x, cdf, px);
  return cdf + px;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  bf162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
