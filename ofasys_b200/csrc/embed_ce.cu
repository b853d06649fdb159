// Adaptor hook (embedding gather / dense rows + position + type -> LayerNorm) and the sum-CE
// criterion kernels.  Both are HBM-bound row kernels: one 128-thread block owns a row of d <= 1024
// (embed) or one 256-thread block a row of V logits (CE).
#include "common.cuh"

namespace {

constexpr int PARTIAL_ROWS = 296;  // == ofab_ln_partial_rows()

__device__ __forceinline__ float block_sum128(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

struct EmbedArgs {
  int B, T, d;
  const int64_t* tokens;
  const bf16 *E, *dense, *cls, *pos, *type, *gamma, *beta;
  int has_cls;
  const uint8_t* zero_mask;
  float eps;
  int drop_on;   // adaptor dropout on the LayerNorm output (adaptor/base.py:181)
  DropArgs drop;
};

// pre-LN value of this thread's 8 columns of row (b, t)
__device__ __forceinline__ f8 embed_pre(const EmbedArgs& a, int b, int t, int c, int64_t& tok_out) {
  f8 v;
  tok_out = -1;
  if (a.tokens != nullptr) {
    const int64_t tok = a.tokens[(int64_t)b * a.T + t];
    tok_out = tok;
    v = load8(a.E + tok * a.d + c);
  } else if (a.has_cls && t == 0) {
    v = load8(a.cls + c);
  } else {
    v = load8(a.dense + ((int64_t)b * (a.T - a.has_cls) + (t - a.has_cls)) * a.d + c);
  }
  if (a.pos != nullptr) {
    const f8 p = load8(a.pos + (int64_t)t * a.d + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] += p.v[j];
  }
  if (a.type != nullptr) {
    const f8 p = load8(a.type + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] += p.v[j];
  }
  return v;
}

__global__ void __launch_bounds__(128) embed_ln_fwd_kernel(const EmbedArgs a, float* __restrict__ out, int64_t out_bs,
                                                           float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ float red[4];
  const int c = threadIdx.x * 8;
  const bool col_ok = c < a.d;
  const float inv_n = 1.0f / (float)a.d;
  f8 g, be;
  if (col_ok) {
    g = load8(a.gamma + c);
    be = load8(a.beta + c);
  }
  const int64_t rows = (int64_t)a.B * a.T;
  DropCtx dk;
  if (a.drop_on) dk = drop_ctx(a.drop);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / a.T), t = (int)(row % a.T);
    f8 v;
    float s = 0.f;
    int64_t tok;
    if (col_ok) {
      v = embed_pre(a, b, t, c, tok);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v.v[j];
    }
    const float mu = block_sum128(s, red) * inv_n;
    float q = 0.f;
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v.v[j] - mu;
        q += dlt * dlt;
      }
    }
    const float rs = rsqrtf(block_sum128(q, red) * inv_n + a.eps);
    if (threadIdx.x == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
    const bool zero = a.zero_mask != nullptr && a.zero_mask[row];
    if (col_ok) {
      f8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = zero ? 0.f : (v.v[j] - mu) * rs * g.v[j] + be.v[j];
      if (a.drop_on) {
        const f8 m = drop_mask8(a.drop, dk, row, c);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] *= m.v[j];
      }
      store8(out + (int64_t)b * out_bs + (int64_t)t * a.d + c, o);
    }
  }
}

// 512 threads = 4 row slots of 128 threads: each slot walks its own rows (named barrier per slot), so the persistent
// grid of PARTIAL_ROWS CTAs keeps 16 warps per CTA in flight; the slots' partial sums are combined at the end.
__device__ __forceinline__ float slot_sum128(float v, float* red /* [4] of this slot */, int slot) {
  v = warp_sum(v);
  asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");
  if ((threadIdx.x & 31) == 0) red[(threadIdx.x >> 5) & 3] = v;
  asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");
  return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(512) embed_ln_bwd_kernel(const EmbedArgs a, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ dout, int64_t dout_bs, float* __restrict__ dE,
                                                           int64_t padding_idx, bf16* __restrict__ ddense, float* __restrict__ dpos,
                                                           float* __restrict__ partial) {
  __shared__ float red_all[4][4];
  __shared__ float comb[3][1024];  // slots 1..3 hand their per-column partial sums to slot 0
  const int slot = threadIdx.x >> 7;
  float* red = red_all[slot];
  const int c = (threadIdx.x & 127) * 8;
  const bool col_ok = c < a.d;
  const float inv_n = 1.0f / (float)a.d;
  f8 g, acc[4];  // dgamma, dbeta, dtype, dcls
  if (col_ok) g = load8(a.gamma + c);
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[s].v[j] = 0.f;
  const int64_t rows = (int64_t)a.B * a.T;
  DropCtx dk;
  if (a.drop_on) dk = drop_ctx(a.drop);
  for (int64_t row = (int64_t)blockIdx.x * 4 + slot; row < rows; row += (int64_t)gridDim.x * 4) {
    const int b = (int)(row / a.T), t = (int)(row % a.T);
    const bool zero = a.zero_mask != nullptr && a.zero_mask[row];
    const float mu = mean[row], rs = rstd[row];
    f8 xh, dy;
    float s1 = 0.f, s2 = 0.f;
    int64_t tok = -1;
    if (col_ok) {
      const f8 pre = embed_pre(a, b, t, c, tok);
      if (zero) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dy.v[j] = 0.f;
      } else {
        dy = load8(dout + (int64_t)b * dout_bs + (int64_t)t * a.d + c);
        if (a.drop_on) {
          const f8 m = drop_mask8(a.drop, dk, row, c);
#pragma unroll
          for (int j = 0; j < 8; ++j) dy.v[j] *= m.v[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh.v[j] = (pre.v[j] - mu) * rs;
        acc[0].v[j] += dy.v[j] * xh.v[j];
        acc[1].v[j] += dy.v[j];
        dy.v[j] *= g.v[j];
        s1 += dy.v[j] * xh.v[j];
        s2 += dy.v[j];
      }
    }
    const float c1 = slot_sum128(s1, red, slot) * inv_n;
    const float c2 = slot_sum128(s2, red, slot) * inv_n;
    if (col_ok) {
      f8 dp;
#pragma unroll
      for (int j = 0; j < 8; ++j) dp.v[j] = rs * (dy.v[j] - c2 - xh.v[j] * c1);
      if (a.type != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[2].v[j] += dp.v[j];
      }
      if (a.pos != nullptr && dpos != nullptr && !zero) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dpos + (int64_t)t * a.d + c + j, dp.v[j]);
      }
      if (a.tokens != nullptr) {
        if (dE != nullptr && tok != padding_idx && !zero) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dE + tok * a.d + c + j, dp.v[j]);
        }
      } else if (a.has_cls && t == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[3].v[j] += dp.v[j];
      } else if (ddense != nullptr) {
        store8(ddense + ((int64_t)b * (a.T - a.has_cls) + (t - a.has_cls)) * a.d + c, dp);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    __syncthreads();
    if (slot > 0 && col_ok) store8(comb[slot - 1] + c, acc[s]);
    __syncthreads();
    if (slot == 0 && col_ok) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const f8 o = load8(comb[k] + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[s].v[j] += o.v[j];
      }
      store8(partial + ((int64_t)s * PARTIAL_ROWS + blockIdx.x) * a.d + c, acc[s]);
    }
  }
}

// ------------------------------------------------------------------------------------------ CE
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; s = 0.f; return; }
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

// eps > 0: label smoothing (label_smoothed_cross_entropy.py:62-92): the row loss becomes
//   (1 - eps - eps_i) * (lse - x_t) + eps_i * (V * lse - sum_c x_c),  eps_i = eps / (V - 1)
// which needs the plain sum of the row's logits next to the online log-sum-exp.
__global__ void __launch_bounds__(256) ce_fwd_kernel(const bf16* __restrict__ logits, int64_t V, int64_t ld, const int64_t* __restrict__ target,
                                                     int64_t ignore_index, float* __restrict__ lse, float* __restrict__ loss_sum, float eps,
                                                     float* __restrict__ nll_sum) {
  __shared__ float sm_m[8], sm_s[8], sm_x[8];
  const int64_t r = blockIdx.x;
  const bf16* row = logits + r * ld;
  float m = -INFINITY, s = 0.f, xs = 0.f;
  const int64_t V8 = V / 8;
  for (int64_t i = threadIdx.x; i < V8; i += 256) {
    const f8 x = load8(row + i * 8);
    float mx = x.v[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) mx = fmaxf(mx, x.v[j]);
    const float mn = fmaxf(m, mx);
    float acc = s * __expf(m - mn);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc += __expf(x.v[j] - mn);
      xs += x.v[j];
    }
    m = mn;
    s = acc;
  }
  for (int64_t i = V8 * 8 + threadIdx.x; i < V; i += 256) {
    const float x = __bfloat162float(row[i]);
    online_merge(m, s, x, 1.f);
    xs += x;
  }
  xs = warp_sum(xs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) {
    sm_m[threadIdx.x >> 5] = m;
    sm_s[threadIdx.x >> 5] = s;
    sm_x[threadIdx.x >> 5] = xs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm_m[0], S = sm_s[0], X = sm_x[0];
    for (int w = 1; w < 8; ++w) {
      online_merge(M, S, sm_m[w], sm_s[w]);
      X += sm_x[w];
    }
    const float l = M + logf(S);
    lse[r] = l;
    const int64_t tg = target[r];
    if (tg != ignore_index) {
      const float nll = l - __bfloat162float(row[tg]);
      float loss = nll;
      if (eps > 0.f) {
        const float eps_i = eps / (float)(V - 1);
        loss = (1.0f - eps - eps_i) * nll + eps_i * ((float)V * l - X);
      }
      atomicAdd(loss_sum, loss);
      if (nll_sum != nullptr) atomicAdd(nll_sum, nll);
    }
  }
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const bf16* __restrict__ logits, int64_t V, int64_t ld, const int64_t* __restrict__ target,
                                                     int64_t ignore_index, const float* __restrict__ lse, const float* __restrict__ gscale,
                                                     bf16* __restrict__ dlogits, float eps) {
  // d loss_row / d x_c = softmax_c - (1 - eps - eps_i) * [c == target] - eps_i      (eps = 0: softmax - onehot)
  const float eps_i = eps > 0.f ? eps / (float)(V - 1) : 0.f;
  const float w_t = 1.0f - eps - eps_i;
  const int64_t r = blockIdx.x;
  const bf16* row = logits + r * ld;
  bf16* drow = dlogits + r * ld;
  const int64_t tg = target[r];
  const bool counted = tg != ignore_index;
  const float g = counted ? gscale[0] : 0.f;
  const float l = lse[r];
  const int64_t V8 = V / 8;
  for (int64_t i = threadIdx.x; i < V8; i += 256) {
    f8 x = load8(row + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float pr = counted ? __expf(x.v[j] - l) : 0.f;
      x.v[j] = g * (pr - ((i * 8 + j) == tg ? w_t : 0.f) - eps_i);
    }
    store8(drow + i * 8, x);
  }
  for (int64_t i = V8 * 8 + threadIdx.x; i < V; i += 256) {
    const float pr = counted ? __expf(__bfloat162float(row[i]) - l) : 0.f;
    drow[i] = __float2bfloat16(g * (pr - (i == tg ? w_t : 0.f) - eps_i));
  }
}

// ---- per-row criterion with constraint masks (label_smoothed_cross_entropy.py:62-92,147-191) ----------------------------------
// Allowed entries of row r: cmask[r, c] != 0 (when given) AND (c < 4 or c_lo <= c < c_hi) (when c_lo >= 0: `constraint_range`,
// get_constraint_masks :147-157); disallowed logits count as -inf.  Per counted row (target != ignore):
//   nll = lse_allowed - x_t;   smooth = sum_{allowed} (lse - x_c);   eps_i = eps / (n_allowed - 1 + 1e-6)  [constraints] or eps / (V - 1)
//   loss = (1 - eps - eps_i) * nll + eps_i * smooth
// The per-row losses go to row_loss / row_nll (the caller applies drop_worst = a top-k over them and sums); the backward takes
// the per-row weight d total / d row_loss.
struct CeRows {
  const bf16* logits;
  int64_t V, ld;
  const int64_t* target;
  int64_t ignore_index;
  const uint8_t* cmask;
  int64_t c_lo, c_hi;
  float eps;
  float *lse, *row_loss, *row_nll, *n_allowed;
  const float* row_scale;
  bf16* dlogits;
};
__device__ __forceinline__ bool ce_allowed(const CeRows& a, int64_t r, int64_t c) {
  bool ok = a.cmask == nullptr || a.cmask[r * a.V + c] != 0;
  if (a.c_lo >= 0) ok = ok && (c < 4 || (c >= a.c_lo && c < a.c_hi));
  return ok;
}
__global__ void __launch_bounds__(256) ce_rows_fwd_kernel(const CeRows a) {
  __shared__ float sm_m[8], sm_s[8], sm_x[8], sm_n[8];
  const int64_t r = blockIdx.x;
  const bf16* row = a.logits + r * a.ld;
  float m = -INFINITY, s = 0.f, xs = 0.f, na = 0.f;
  for (int64_t i = threadIdx.x; i < a.V; i += 256) {
    if (!ce_allowed(a, r, i)) continue;
    const float x = __bfloat162float(row[i]);
    online_merge(m, s, x, 1.f);
    xs += x;
    na += 1.f;
  }
  xs = warp_sum(xs);
  na = warp_sum(na);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) {
    sm_m[threadIdx.x >> 5] = m; sm_s[threadIdx.x >> 5] = s; sm_x[threadIdx.x >> 5] = xs; sm_n[threadIdx.x >> 5] = na;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm_m[0], S = sm_s[0], X = sm_x[0], N = sm_n[0];
    for (int w = 1; w < 8; ++w) {
      online_merge(M, S, sm_m[w], sm_s[w]);
      X += sm_x[w];
      N += sm_n[w];
    }
    const float l = M + logf(S);
    a.lse[r] = l;
    a.n_allowed[r] = N;
    const int64_t tg = a.target[r];
    float loss = 0.f, nll = 0.f;
    if (tg != a.ignore_index) {
      nll = l - __bfloat162float(row[tg]);  // (a target outside the allowed set gives +inf, as the reference's gather on -inf lprobs)
      if (!ce_allowed(a, r, tg)) nll = INFINITY;
      loss = nll;
      if (a.eps > 0.f) {
        const bool constrained = a.cmask != nullptr || a.c_lo >= 0;
        const float eps_i = constrained ? a.eps / (N - 1.0f + 1e-6f) : a.eps / (float)(a.V - 1);
        loss = (1.0f - a.eps - eps_i) * nll + eps_i * (N * l - X);
      }
    }
    a.row_loss[r] = loss;
    a.row_nll[r] = nll;
  }
}
__global__ void __launch_bounds__(256) ce_rows_bwd_kernel(const CeRows a) {
  // d row_loss / d x_c = softmax_c - (1 - eps - eps_i) [c == target] - eps_i   on the allowed entries, 0 elsewhere
  const int64_t r = blockIdx.x;
  const bf16* row = a.logits + r * a.ld;
  bf16* drow = a.dlogits + r * a.ld;
  const int64_t tg = a.target[r];
  const bool counted = tg != a.ignore_index;
  const float g = counted ? a.row_scale[r] : 0.f;
  const float l = a.lse[r];
  const bool constrained = a.cmask != nullptr || a.c_lo >= 0;
  const float eps_i = a.eps > 0.f ? (constrained ? a.eps / (a.n_allowed[r] - 1.0f + 1e-6f) : a.eps / (float)(a.V - 1)) : 0.f;
  const float w_t = 1.0f - a.eps - eps_i;
  for (int64_t i = threadIdx.x; i < a.ld; i += 256) {
    float v = 0.f;
    if (i < a.V && g != 0.f && ce_allowed(a, r, i)) v = g * (__expf(__bfloat162float(row[i]) - l) - (i == tg ? w_t : 0.f) - eps_i);
    drow[i] = __float2bfloat16(v);
  }
}

int to_args(const ofab_embed_ln_args* a, EmbedArgs& e, const char* who) {
  OFAB_REQUIRE(a->B > 0 && a->T > 0 && a->d >= 8 && a->d <= 1024 && a->d % 8 == 0, "%s: bad shape B=%d T=%d d=%d (d multiple of 8, <= 1024)", who, a->B, a->T, a->d);
  OFAB_REQUIRE((a->tokens != nullptr) != (a->dense != nullptr || (a->has_cls && a->T == 1)), "%s: exactly one of tokens / dense must be given", who);
  OFAB_REQUIRE(a->tokens == nullptr || a->E != nullptr, "%s: E is NULL", who);
  OFAB_REQUIRE(!a->has_cls || a->cls != nullptr, "%s: has_cls without cls", who);
  OFAB_REQUIRE(a->gamma && a->beta && a->mean && a->rstd, "%s: gamma/beta/mean/rstd NULL", who);
  e.B = a->B; e.T = a->T; e.d = a->d;
  e.tokens = a->tokens;
  e.E = (const bf16*)a->E; e.dense = (const bf16*)a->dense; e.cls = (const bf16*)a->cls; e.pos = (const bf16*)a->pos;
  e.type = (const bf16*)a->type; e.gamma = (const bf16*)a->gamma; e.beta = (const bf16*)a->beta;
  e.has_cls = a->has_cls ? 1 : 0;
  e.zero_mask = a->zero_mask;
  e.eps = a->eps;
  e.drop_on = a->drop != nullptr && (a->drop->p > 0.f || a->drop->drop_path > 0.f);
  e.drop = DropArgs{};
  if (e.drop_on && !ofab_drop_args(a->drop, e.drop, who)) return OFAB_ERR_ARG;
  return OFAB_OK;
}

}  // namespace

extern "C" int ofab_embed_ln_fwd(const ofab_embed_ln_args* a, ofab_stream_t stream) {
  EmbedArgs e;
  int rc = to_args(a, e, "ofab_embed_ln_fwd");
  if (rc) return rc;
  OFAB_REQUIRE(a->out != nullptr && a->out_bs >= (int64_t)a->T * a->d && a->out_bs % 4 == 0, "ofab_embed_ln_fwd: bad out / out_bs");
  const int64_t rows = (int64_t)a->B * a->T;
  const int64_t cap = (int64_t)ofab_sm_count() * 16;
  embed_ln_fwd_kernel<<<(unsigned)(rows < cap ? rows : cap), 128, 0, (cudaStream_t)stream>>>(e, a->out, a->out_bs, a->mean, a->rstd);
  OFAB_LAUNCH_CHECK("ofab_embed_ln_fwd");
  return OFAB_OK;
}

extern "C" int ofab_embed_ln_bwd(const ofab_embed_ln_bwd_args* a, ofab_stream_t stream) {
  EmbedArgs e;
  int rc = to_args(&a->f, e, "ofab_embed_ln_bwd");
  if (rc) return rc;
  OFAB_REQUIRE(a->dout != nullptr && a->dgb_partial != nullptr, "ofab_embed_ln_bwd: dout / dgb_partial NULL");
  embed_ln_bwd_kernel<<<PARTIAL_ROWS, 512, 0, (cudaStream_t)stream>>>(e, a->f.mean, a->f.rstd, a->dout, a->dout_bs, a->dE, a->padding_idx,
                                                                      (bf16*)a->ddense, a->dpos, a->dgb_partial);
  OFAB_LAUNCH_CHECK("ofab_embed_ln_bwd");
  return OFAB_OK;
}

extern "C" int ofab_ce_fwd(const void* logits, int64_t rows, int64_t V, int64_t ld, const int64_t* target, int64_t ignore_index,
                           float* lse, float* loss_sum, float label_smoothing, float* nll_sum, ofab_stream_t stream) {
  OFAB_REQUIRE(rows > 0 && V > 1 && ld >= V && ld % 8 == 0, "ofab_ce_fwd: bad shape rows=%lld V=%lld ld=%lld (ld multiple of 8)", (long long)rows, (long long)V, (long long)ld);
  OFAB_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, "ofab_ce_fwd: label_smoothing=%g out of [0, 1)", (double)label_smoothing);
  ce_fwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const bf16*)logits, V, ld, target, ignore_index, lse, loss_sum, label_smoothing, nll_sum);
  OFAB_LAUNCH_CHECK("ofab_ce_fwd");
  return OFAB_OK;
}

static int ce_rows_args(const ofab_ce_rows_args* a, CeRows& c, const char* who) {
  OFAB_REQUIRE(a->rows > 0 && a->V > 1 && a->ld >= a->V && a->ld % 8 == 0, "%s: bad shape rows=%lld V=%lld ld=%lld", who, (long long)a->rows, (long long)a->V, (long long)a->ld);
  OFAB_REQUIRE(a->label_smoothing >= 0.f && a->label_smoothing < 1.f, "%s: label_smoothing=%g out of [0, 1)", who, (double)a->label_smoothing);
  OFAB_REQUIRE(a->logits && a->target && a->lse && a->n_allowed, "%s: NULL tensor", who);
  OFAB_REQUIRE(a->c_lo < 0 || (a->c_lo >= 4 && a->c_hi > a->c_lo && a->c_hi <= a->V), "%s: constraint range [%lld, %lld) outside [4, V]", who, (long long)a->c_lo, (long long)a->c_hi);
  c.logits = (const bf16*)a->logits; c.V = a->V; c.ld = a->ld; c.target = a->target; c.ignore_index = a->ignore_index;
  c.cmask = a->cmask; c.c_lo = a->c_lo; c.c_hi = a->c_hi; c.eps = a->label_smoothing;
  c.lse = a->lse; c.row_loss = a->row_loss; c.row_nll = a->row_nll; c.n_allowed = a->n_allowed;
  c.row_scale = a->row_scale; c.dlogits = (bf16*)a->dlogits;
  return OFAB_OK;
}
extern "C" int ofab_ce_rows_fwd(const ofab_ce_rows_args* a, ofab_stream_t stream) {
  CeRows c;
  int rc = ce_rows_args(a, c, "ofab_ce_rows_fwd");
  if (rc) return rc;
  OFAB_REQUIRE(a->row_loss && a->row_nll, "ofab_ce_rows_fwd: row_loss / row_nll NULL");
  ce_rows_fwd_kernel<<<(unsigned)a->rows, 256, 0, (cudaStream_t)stream>>>(c);
  OFAB_LAUNCH_CHECK("ofab_ce_rows_fwd");
  return OFAB_OK;
}
extern "C" int ofab_ce_rows_bwd(const ofab_ce_rows_args* a, ofab_stream_t stream) {
  CeRows c;
  int rc = ce_rows_args(a, c, "ofab_ce_rows_bwd");
  if (rc) return rc;
  OFAB_REQUIRE(a->row_scale && a->dlogits, "ofab_ce_rows_bwd: row_scale / dlogits NULL");
  ce_rows_bwd_kernel<<<(unsigned)a->rows, 256, 0, (cudaStream_t)stream>>>(c);
  OFAB_LAUNCH_CHECK("ofab_ce_rows_bwd");
  return OFAB_OK;
}

extern "C" int ofab_ce_bwd(const void* logits, int64_t rows, int64_t V, int64_t ld, const int64_t* target, int64_t ignore_index,
                           const float* lse, const float* gscale, void* dlogits, float label_smoothing, ofab_stream_t stream) {
  OFAB_REQUIRE(rows > 0 && V > 1 && ld >= V && ld % 8 == 0, "ofab_ce_bwd: bad shape");
  OFAB_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, "ofab_ce_bwd: label_smoothing=%g out of [0, 1)", (double)label_smoothing);
  ce_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const bf16*)logits, V, ld, target, ignore_index, lse, gscale, (bf16*)dlogits,
                                                                  label_smoothing);
  OFAB_LAUNCH_CHECK("ofab_ce_bwd");
  return OFAB_OK;
}
