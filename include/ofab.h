/*
 * ofab.h -- C ABI of the B200-native OFASys hot path (libofab.so, sm_100a only).
 *
 * The reference (OFA-Sys/OFASys) has no C ABI for this path: its model is Python classes found
 * through the ConfigStore registry (ofasys/model/ofa.py:328, ofasys/adaptor/general.py:69-93) and
 * its only native precedent is the pybind11 module pair `scaled_softmax_cuda.{forward,backward}` /
 * `scaled_masked_softmax_cuda.*` (ofasys/module/fused_kernels/scaled_softmax.cpp:67-74,
 * scaled_masked_softmax.cpp:84-97), which take torch::Tensor.  This header is what a maintainer
 * binds instead (ctypes stub in INTEGRATION.md): plain device pointers, sizes and a stream.
 *
 * Conventions
 *  - every entry point returns 0 on success, a negative code on error (never throws);
 *    `ofab_last_error()` returns a thread-local message.
 *  - no allocation inside: the caller owns every buffer (outputs, saved statistics, workspaces).
 *  - all launches go to the caller's stream; entry points are stateless and re-entrant per stream.
 *  - activations are row-major [rows, cols] with rows = batch*time ("token rows").
 *  - dt arguments: 0 = float32, 1 = bfloat16.
 */
#ifndef OFAB_H_
#define OFAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ofab_stream_t; /* cudaStream_t */

#define OFAB_F32 0
#define OFAB_BF16 1

#define OFAB_OK 0
#define OFAB_ERR_ARG -1     /* bad argument (shape/alignment/dtype) */
#define OFAB_ERR_CUDA -2    /* CUDA runtime / driver error */
#define OFAB_ERR_DEVICE -3  /* not an sm_100 device */

int ofab_version(void);
const char* ofab_last_error(void);
/* 0 if `device` is compute capability 10.x; OFAB_ERR_DEVICE otherwise (replaces the reference's
 * arch list sm_70/sm_80, ofasys/module/fused_kernels/__init__.py:29-47). */
int ofab_device_check(int device);
int ofab_num_sms(void);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm family.  Replaces torch.nn.LayerNorm as built by ofasys/module/layer_norm.py:27-32
 * (eps 1e-5) and the never-built apex kernels cuApplyLayerNorm / cuComputeGradInput
 * (ofasys/module/fused_kernels/layer_norm_cuda_kernel.cu:290-334,534).
 * Statistics are fp32; gamma/beta are bf16 (reference `common.bf16` casts LN params too,
 * ofasys/engine/trainer.py:215-219).
 * ------------------------------------------------------------------------------------------- */

/* y = LN(act(x)) ; act = exact-erf GELU in fp32 when `gelu` != 0 (ofasys/module/gelu.py:18-19 +
 * ffn_layernorm, ofasys/module/transformer_layer.py:188-196), identity otherwise.
 * x: [rows, cols] x_dt; y: [rows, cols] y_dt; mean/rstd: [rows] fp32 (saved for backward). */
int ofab_ln_fwd(const void* x, int x_dt, const void* gamma, const void* beta, void* y, int y_dt,
                float* mean, float* rstd, int64_t rows, int cols, float eps, int gelu,
                ofab_stream_t stream);

/* backward of ofab_ln_fwd.  dx (dx_dt) = d act(x) * LN'(dy); if `dx_accum` != 0, dx += (fp32 only).
 * dgb_partial: fp32 [3, ofab_ln_partial_rows(), cols] scratch that receives per-block partial sums
 * of dgamma (slab 0), dbeta (slab 1) and of dx itself (slab 2: column sums of dx = the bias gradient of
 * the Linear layer that produced x, nn.Linear backward); finish with ofab_reduce_partials. */
int ofab_ln_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const void* gamma,
                const float* mean, const float* rstd, void* dx, int dx_dt, int dx_accum,
                float* dgb_partial, int64_t rows, int cols, int gelu, ofab_stream_t stream);
int ofab_ln_partial_rows(void);

/* Fused "normformer" junction of every attention block (transformer_layer.py:175-186,428-436):
 *   a_ln = LN1(a);  x_new = x + a_ln;  y = LN2(x_new)
 * a: bf16 [rows, cols]; x, x_new: fp32; y: bf16.  stats: fp32 [4, rows] = mean1,rstd1,mean2,rstd2.
 * g1 == b1 == NULL: no first LayerNorm, x_new = x + a (the FFN's deferred residual add fused with the next
 * block's pre-LayerNorm, transformer_layer.py:203-209 + :170); dg1/db1 partial slabs are then zero. */
int ofab_ln_res_ln_fwd(const void* a, const float* x, const void* g1, const void* b1, const void* g2,
                       const void* b2, float* x_new, void* y, float* stats, int64_t rows, int cols,
                       float eps, ofab_stream_t stream);
/* backward: dx_tot = dx_new + LN2'(dy);  da = LN1'(dx_tot).  dx_new may alias dx_tot.
 * dgb_partial: fp32 [5, ofab_ln_partial_rows(), cols] = dg1, db1, dg2, db2 partials and the column sums of
 * da (bias gradient of the Linear that produced a). */
int ofab_ln_res_ln_bwd(const float* dx_new, const void* dy, const void* a, const float* x_new,
                       const void* g1, const void* g2, const float* stats, float* dx_tot, void* da,
                       float* dgb_partial, int64_t rows, int cols, ofab_stream_t stream);

/* out[c] (+)= sum_r in[r, c].  in: [rows, cols] in_dt (fp32 partials or bf16 activations grads);
 * out: [cols] out_dt.  Used for LN dgamma/dbeta partials, Linear bias grads, batch sums.
 * Deterministic (two-stage, no atomics). */
int ofab_colsum(const void* in, int in_dt, int64_t rows, int64_t cols, int64_t ld, void* out,
                int out_dt, int accumulate, float* scratch, ofab_stream_t stream);
/* out[s, c] = sum over the ofab_ln_partial_rows() rows of slab s of partial[nslabs, rows, cols]:
 * finishes dgamma / dbeta (and dtype / dcls) of one backward kernel in a single launch. */
int ofab_reduce_partials(const float* partial, int nslabs, int cols, void* out, int out_dt,
                         ofab_stream_t stream);
/* floats of `scratch` ofab_colsum needs for `cols` columns */
int64_t ofab_colsum_scratch_elems(int64_t cols);

/* ---------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores (TMA -> smem ring -> tcgen05.mma, fp32 accumulators in TMEM).
 * Replaces every F.linear / addmm on the path (multihead_attention.py:199-218,346;
 * transformer_layer.py:188-203; adaptor/base.py:131 tied logits) and their autograd backward.
 *   D[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ residual[M,N])
 * A and B are bf16.  `a_mn_major` = 0: A stored row-major [M,K] (K contiguous, "K-major");
 * 1: A stored [K,M] (M contiguous) -- i.e. the transposed operand is read in place (wgrad).
 * Same for B with N.  lda/ldb/ldd/ldr are leading dimensions in elements (multiples of 8).
 * bias: bf16 [N] or NULL.  residual: fp32 [M, ldr] or NULL.  D: d_dt (fp32 or bf16).
 * ------------------------------------------------------------------------------------------- */
int ofab_gemm_bf16(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, int a_mn_major,
                   const void* B, int64_t ldb, int b_mn_major, const void* bias,
                   const float* residual, int64_t ldr, void* D, int64_t ldd, int d_dt,
                   ofab_stream_t stream);

/* Split-K form for outputs with too few tiles to fill the chip and a long contraction -- the weight gradients
 * dW[N_out, K_in] = dY^T X of nn.Linear backward (autograd of the F.linear calls above), whose tile count does not
 * grow with the batch while K = B*T does.  D = A * B^T (no bias / residual); the K loop is cut into ranges that run
 * as separate work items writing fp32 partial slabs into `workspace`, then one reduction pass writes D (bf16/fp32).
 * The number of ranges is chosen internally; ofab_gemm_splitk_workspace_elems() returns the fp32 elements of
 * workspace this call needs for (M, N, K) -- 0 means the plain kernel is used and workspace may be NULL. */
int64_t ofab_gemm_splitk_workspace_elems(int64_t M, int64_t N, int64_t K);
int ofab_gemm_bf16_splitk(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, int a_mn_major,
                          const void* B, int64_t ldb, int b_mn_major, void* D, int64_t ldd, int d_dt,
                          float* workspace, int64_t workspace_elems, ofab_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused attention (flash-style; scores, bias and probabilities never touch HBM).
 * Replaces MultiheadAttention.forward's bmm/+bias/mask/softmax/bmm chain
 * (ofasys/module/multihead_attention.py:308-338) and the vendored Megatron softmax kernels
 * scaled_softmax_warp_forward / scaled_masked_softmax_warp_{forward,backward}
 * (ofasys/module/fused_kernels/scaled_masked_softmax.h:99-423).
 *
 *   S[b,h,i,j] = scale * ( q[b,i,h,:] . k[b,j,h,:]  +  pq[b,i,h,:] . pk[b,j,h,:] )
 *              + table[rp_idx[i,j], h]            (rp_idx >= 0)
 *              ; -inf where (causal and j > i) or kpm[b,j]
 *   P = softmax_j(S) in fp32 ;  o[b,i,h,:] = sum_j P v[b,j,h,:]
 * head_dim is 64 (every OFA preset: ofasys/model/ofa.py:557-650).  q/k/v/o are bf16 with element
 * strides (batch, row); head h occupies columns [h*64, h*64+64).  pq/pk (bf16) carry the absolute
 * position terms of OFAGeneralAdaptor.build_abs_pos_bias (ofasys/adaptor/general.py:223-243) /
 * TransformerDecoder.get_cross_pos_info (ofasys/model/transformer.py:280-299) as extra
 * contraction columns; NULL when absent; a batch stride of 0 broadcasts them.
 * rp_idx: int32 [Tq, Tk] bucket ids (the adaptor's *_rp_bucket gathers, adaptor/text.py:101-104,
 * image_resnet.py:116-128) or NULL; table: fp32 [n_buckets, H].  kpm: uint8 [B, Tk] or NULL.
 * lse: fp32 [B, H, Tq] log-sum-exp saved for backward.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int B, H, Tq, Tk;
  const void *q, *k, *v;     /* bf16 */
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs;
  const void *pq, *pk;       /* bf16 or NULL */
  int64_t pq_bs, pq_rs, pk_bs, pk_rs;
  const int32_t* rp_idx;     /* [Tq, Tk] or NULL */
  const float* table;        /* [n_buckets, H] or NULL */
  int n_buckets;
  const uint8_t* kpm;        /* [B, Tk] or NULL */
  int causal;
  float scale;
  void* o;                   /* bf16 [B, Tq, H*64] */
  int64_t o_bs, o_rs;
  float* lse;                /* [B, H, Tq] */
} ofab_attn_fwd_args;

int ofab_attn_fwd(const ofab_attn_fwd_args* args, ofab_stream_t stream);

typedef struct {
  ofab_attn_fwd_args f;      /* same tensors as forward (o, lse = saved outputs) */
  const void* d_o;           /* bf16 [B, Tq, H*64] */
  int64_t do_bs, do_rs;
  void *dq, *dk, *dv;        /* bf16, strides below */
  int64_t dq_bs, dq_rs, dk_bs, dk_rs, dv_bs, dv_rs;
  void *dpq, *dpk;           /* bf16 per-batch grads [B, T, H*64] (contiguous) or NULL */
  float* dtable;             /* fp32 [n_buckets, H], accumulated with atomics (caller zeroes) or NULL */
  float* delta;              /* fp32 scratch [B, H, Tq] */
} ofab_attn_bwd_args;

int ofab_attn_bwd(const ofab_attn_bwd_args* args, ofab_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Adaptor hook: embedding row + position + type embedding -> LayerNorm
 * (ofasys/adaptor/base.py:152-191 with ofasys/adaptor/text.py:106-127 or a dense adaptor output).
 *   pre[b,t,:] = src(b,t) (+ pos[t,:]) (+ type[:]) ;  out = LN(pre) * keep(b,t)
 * src = E[tokens[b,t]] when tokens != NULL (nn.Embedding gather, padding_idx row included as
 * stored), else dense[b, t - has_cls] (bf16) with the cls row for t == 0 when has_cls.
 * keep = 0 for rows where zero_mask[b,t] != 0 (TransformerEncoder.forward zeroes padded rows,
 * ofasys/model/transformer.py:109-112).  out: fp32 with batch stride out_bs (lets adaptors write
 * straight into their slice of the concatenated sequence, ofasys/adaptor/general.py:245-260).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int B, T, d;
  const int64_t* tokens;     /* [B, T] or NULL */
  const void* E;             /* bf16 [V, d] */
  const void* dense;         /* bf16 [B, T - has_cls, d] */
  const void* cls;           /* bf16 [d] or NULL */
  int has_cls;
  const void* pos;           /* bf16 [>=T, d] rows indexed by t, or NULL */
  const void* type;          /* bf16 [d] or NULL */
  const void *gamma, *beta;  /* bf16 [d] */
  const uint8_t* zero_mask;  /* [B, T] or NULL */
  float eps;
  float* out;                /* fp32 */
  int64_t out_bs;            /* elements between batches of out */
  float *mean, *rstd;        /* [B*T] */
} ofab_embed_ln_args;

int ofab_embed_ln_fwd(const ofab_embed_ln_args* a, ofab_stream_t stream);

/* backward: dE fp32 [V,d] (atomic scatter-add, rows == padding_idx skipped), ddense bf16,
 * dpos fp32 [T,d] (atomic), dcls/dtype via dgb_partial slabs 2/3.
 * dgb_partial: fp32 [4, ofab_ln_partial_rows(), d] = dgamma, dbeta, dtype, dcls partials. */
typedef struct {
  ofab_embed_ln_args f;
  const float* dout;         /* fp32, batch stride dout_bs */
  int64_t dout_bs;
  float* dE;                 /* fp32 [V, d] or NULL */
  int64_t padding_idx;       /* -1: none */
  void* ddense;              /* bf16 [B, T-has_cls, d] or NULL */
  float* dpos;               /* fp32 [T, d] or NULL (accumulated) */
  float* dgb_partial;
} ofab_embed_ln_bwd_args;

int ofab_embed_ln_bwd(const ofab_embed_ln_bwd_args* a, ofab_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Criterion boundary: sum-reduced cross entropy over non-pad targets with fp32 log-softmax
 * (ofasys/engine/criterion/cross_entropy.py:27-41,62-67; module/utils.py:458-462).
 * logits: bf16 [rows, ld] (first V columns valid); target: int64 [rows].
 * fwd: lse[r] = logsumexp(logits[r,:V]); loss_sum += sum_{target != ignore} (lse - logit[target]).
 * bwd: dlogits[r, c] = gscale * (softmax - onehot) for counted rows, 0 otherwise (may alias logits).
 * ------------------------------------------------------------------------------------------- */
int ofab_ce_fwd(const void* logits, int64_t rows, int64_t V, int64_t ld, const int64_t* target,
                int64_t ignore_index, float* lse, float* loss_sum, ofab_stream_t stream);
int ofab_ce_bwd(const void* logits, int64_t rows, int64_t V, int64_t ld, const int64_t* target,
                int64_t ignore_index, const float* lse, const float* gscale, void* dlogits,
                ofab_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small data-movement kernels on the path.
 * ------------------------------------------------------------------------------------------- */
/* y = (bf16) x, optionally colsum: bias_grad[c] = sum_r x[r,c] (fp32 [cols] scratch, zeroed by caller). */
int ofab_cast_f32_bf16(const float* x, void* y, int64_t n, ofab_stream_t stream);
int ofab_cast_bf16_f32(const void* x, float* y, int64_t n, ofab_stream_t stream);
/* out = a + b (fp32) */
/* Multi-tensor copy (data-parallel exchange step): `chunks` is a DEVICE array of n_chunks records
 * {const void* src; void* dst; uint64 bytes} (24 bytes each); one launch copies them all.  Used to pack the
 * per-parameter gradients of a bucket into the flat all-reduce buffer and back (replaces the per-bucket flatten /
 * unflatten of c10d DDP's Reducer that the reference relies on: distributed_model_dispatcher.py:49-75). */
int ofab_multi_copy(const void* chunks, int64_t n_chunks, ofab_stream_t stream);
int ofab_add_f32(const float* a, const float* b, float* out, int64_t n, ofab_stream_t stream);
/* W_eff[n, k] = W[n, k] * c[k / group]  (per-head c_attn folded into out_proj,
 * ofasys/module/multihead_attention.py:342-346).  bf16 in/out, c bf16 [cols/group]. */
int ofab_scale_cols(const void* W, const void* c, void* out, int64_t rows, int64_t cols, int group,
                    ofab_stream_t stream);
/* backward of scale_cols: dW = dW_eff * c ; dc[h] = sum dW_eff * W over the head's columns.
 * dc: fp32 [cols/group], zeroed by the caller (atomics). */
int ofab_scale_cols_bwd(const void* dW_eff, const void* W, const void* c, void* dW, float* dc,
                        int64_t rows, int64_t cols, int group, ofab_stream_t stream);
/* im2col for non-overlapping patches (Conv2d k == stride, ofasys/adaptor/image_patch_embed.py:58-66):
 * img [B, C, H, W] (img_dt) -> cols bf16 [B*(H/p)*(W/p), ldk], column = c*p*p + ph*p + pw,
 * columns [C*p*p, ldk) zero-filled. */
int ofab_patch_im2col(const void* img, int img_dt, int B, int C, int H, int W, int p, void* cols,
                      int64_t ldk, ofab_stream_t stream);
/* general im2col for Conv2d(kernel 3, stride 2, no padding) on channel-last activations
 * (ofasys/module/subsample.py:27-31): x bf16 [B, Hin, Win, C] -> cols [B*Hout*Wout, 9*C],
 * column = (kh*3 + kw)*C + c. */
int ofab_im2col_3x3s2(const void* x, int B, int Hin, int Win, int C, void* cols, ofab_stream_t stream);
/* its adjoint: dx[B,Hin,Win,C] (bf16, overwritten) = col2im(dcols) */
int ofab_col2im_3x3s2(const void* dcols, int B, int Hin, int Win, int C, void* dx, ofab_stream_t stream);
/* first audio conv: Conv2d(1, C, 3, stride 2) + ReLU on fbank [B, L, F] (in_dt) -> bf16 [B, H1, W1, C]
 * channel-last; w bf16 [C, 9], bias bf16 [C]. */
int ofab_conv1_relu_fwd(const void* fbank, int in_dt, int B, int L, int F, const void* w, const void* bias,
                        int C, void* out, ofab_stream_t stream);
/* dw fp32 [C, 9], db fp32 [C] (atomics, caller zeroes); dy bf16 [B,H1,W1,C]; y = saved output (ReLU mask). */
int ofab_conv1_relu_bwd(const void* fbank, int in_dt, int B, int L, int F, const void* y, const void* dy,
                        int C, float* dw, float* db, ofab_stream_t stream);
/* out[o, b, a] = in[o, a, b] (bf16): maps the reference's NCHW conv weight [Co, Ci, 3*3] onto the
 * im2col column order [Co, 3*3, Ci] and the subsampler's Linear(C*F', d) weight (feature index
 * c*F' + f, ofasys/module/subsample.py:61) onto channel-last f*C + c.  Its own inverse with A, B swapped. */
int ofab_transpose_last2(const void* in, void* out, int64_t O, int A, int Bdim, ofab_stream_t stream);
/* y = relu(x) in place helpers for the second conv: dy *= (y > 0) */
int ofab_relu_bwd_inplace(const void* y, void* dy, int64_t n, ofab_stream_t stream);
int ofab_relu_inplace(void* y, int64_t n, ofab_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ResNet backbone of the image / video adaptors (ofasys/module/resnet.py:116-246; adaptor/image_resnet.py:
 * 166-202).  Activations are channel-last bf16 [B, H, W, C] == token rows [B*H*W, C]: 1x1 convolutions are
 * ofab_gemm_bf16 directly, k x k convolutions are im2col + ofab_gemm_bf16 (weights permuted [Co,Ci,k*k] ->
 * [Co,k*k,Ci] with ofab_transpose_last2).
 * ------------------------------------------------------------------------------------------- */
/* Video clip [B, C, F, H, W] (f32 or bf16) -> frames bf16 [B*F, C, H, W] (frames may be NULL) and zero[B*F] = 1 when
 * every value of the frame is 0: the frame-padding test of VideoImageSequenceAdaptor.get_clip_videos_info
 * (adaptor/video_image_sequence.py:118-139, `clip.abs().mean(-1) == 0`; bit-exact, NaN counts as non-zero). */
int ofab_video_frames(const void* video, int dt, int B, int C, int F, int64_t HW, void* frames, uint8_t* zero, ofab_stream_t stream);

/* stem: img [B,C,H,W] (img_dt) -> cols bf16 [B*Ho*Wo, ldk], column = c*k*k + i*k + j (the weight's own
 * flattening, resnet.py:167 conv1 7x7 s2 p3), zero padding, columns >= C*k*k zero. */
int ofab_im2col_nchw(const void* img, int img_dt, int B, int C, int H, int W, int k, int stride, int pad,
                     void* cols, int64_t ldk, ofab_stream_t stream);
/* x bf16 [B,H,W,C] -> cols [B*Ho*Wo, k*k*C], column = (i*k + j)*C + c, zero padding (conv3x3, resnet.py:20-32) */
int ofab_im2col_nhwc(const void* x, int B, int H, int W, int C, int k, int stride, int pad, void* cols,
                     ofab_stream_t stream);
/* adjoint of ofab_im2col_nhwc: dx [B,H,W,C] (overwritten) */
int ofab_col2im_nhwc(const void* dcols, int B, int H, int W, int C, int k, int stride, int pad, void* dx,
                     ofab_stream_t stream);
/* spatial stride-2 row selection of a 1x1 stride-2 convolution (downsample, resnet.py:207-210).
 * backward == 0: x [B,H,W,C] -> y [B,ceil(H/2),ceil(W/2),C]; backward != 0: x = dy, y = dx (zero-filled scatter). */
int ofab_subsample2(const void* x, int B, int H, int W, int C, void* y, int backward, ofab_stream_t stream);
/* nn.MaxPool2d(3, stride 2, padding 1) on [B,H,W,C]; argmax: uint8 tap index (first maximum) for backward */
int ofab_maxpool3x3s2_fwd(const void* x, int B, int H, int W, int C, void* y, uint8_t* argmax, ofab_stream_t stream);
int ofab_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, int B, int H, int W, int C, void* dx, ofab_stream_t stream);
/* nn.BatchNorm2d in training mode on x bf16 [R, C] (R = B*H*W): batch mean / biased variance (fp32 out), momentum
 * update of running_mean / running_var (unbiased), run_dt = dtype of the running buffers (NULL to skip).
 * scratch: ofab_bn_scratch_elems(C) floats. */
int64_t ofab_bn_scratch_elems(int C);
int ofab_bn_stats(const void* x, int64_t R, int C, float* mean, float* var, void* run_mean, void* run_var,
                  int run_dt, float momentum, float* scratch, ofab_stream_t stream);
/* y = relu?( (x - mean) * rsqrt(var + eps) * gamma + beta (+ residual) )  -- the bottleneck tail
 * `out = identity + bn3(conv3(..)); relu` (resnet.py:131-134) is one pass. */
int ofab_bn_apply(const void* x, const float* mean, const float* var, const void* gamma, const void* beta,
                  const void* residual, void* y, int64_t R, int C, float eps, int relu, ofab_stream_t stream);
/* backward of stats+apply: g = dy * (y > 0 if relu); sums[0..C) = sum g (= dbeta), sums[C..2C) = sum g*xhat
 * (= dgamma); dx = gamma*rstd*(g - sum_g/R - xhat*sum_gxhat/R); dres (optional) = g. */
int ofab_bn_bwd(const void* dy, const void* x, const void* y, const float* mean, const float* var,
                const void* gamma, float* sums, void* dx, void* dres, int64_t R, int C, float eps, int relu,
                float* scratch, ofab_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OFAB_H_ */
