"""torch.autograd.Function wrappers over the C ABI (include/ofab.h).

PyTorch supplies device memory (caching allocator), the current CUDA stream and autograd's graph
walk; every kernel that runs is ours.  All functions require CUDA tensors and raise otherwise --
there is no CPU or eager fallback.
"""
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib import BF16, F32

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.OfabError("ofasys_b200 ops need CUDA tensors (no CPU fallback); got a CPU tensor")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------ dropout state
class DropoutState:
    """Seed + step counter of the counter-based dropout masks (include/ofab.h `ofab_dropout`), one per device.

    The two int64 words live in DEVICE memory and are read by the kernels when they run, so a captured CUDA graph
    draws fresh masks on every replay: `next_step()` is a device-side increment that is part of the captured work.
    Call sites of one step are told apart by `site`, a host counter restarted by next_step(); backward kernels get
    the descriptor their forward used and regenerate the same mask (no mask tensor is ever stored).
    """

    def __init__(self, device, seed=None):
        seed = torch.initial_seed() if seed is None else seed
        self.state = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        self._one = torch.ones(1, dtype=torch.int64, device=device)
        self.site = 0

    def reseed(self, seed, step=0):
        self.state.copy_(torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, step], dtype=torch.int64))
        self.site = 0

    def next_step(self):
        self.state[1:2].add_(self._one)
        self.site = 0

    def spec(self, p=0.0, drop_path=0.0, rows_per_sample=0):
        """A fresh call-site descriptor, or None when nothing would be dropped."""
        if p <= 0.0 and drop_path <= 0.0:
            return None
        self.site += 1
        d = _lib.Dropout()
        d.state, d.site, d.p, d.drop_path, d.rows_per_sample = self.state.data_ptr(), self.site, float(p), float(drop_path), int(rows_per_sample)
        d._keep = self.state  # the descriptor holds a raw pointer into it
        return d


_DROPOUT_STATES = {}


def dropout_state(device=None) -> DropoutState:
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    st = _DROPOUT_STATES.get(device)
    if st is None:
        st = _DROPOUT_STATES[device] = DropoutState(device)
    return st


def _dp(drop):
    return None if drop is None else ctypes.byref(drop)


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, drop):
        _need_cuda(x)
        x = _c(x)
        cols = x.shape[-1]
        y = torch.empty_like(x)
        _lib.call("ofab_dropout_apply", _p(x), _p(y), _DT[x.dtype], x.numel() // cols, cols, _dp(drop), _s())
        ctx.drop = drop
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        cols = dy.shape[-1]
        dx = torch.empty_like(dy)
        _lib.call("ofab_dropout_apply", _p(dy), _p(dx), _DT[dy.dtype], dy.numel() // cols, cols, _dp(ctx.drop), _s())
        return dx, None


def dropout(x, drop):
    """x * keep-mask of descriptor `drop` (DropoutState.spec); identity when drop is None.  Rows = all leading dims."""
    return x if drop is None else _DropoutFn.apply(x, drop)


def dropout_mask(drop, rows, cols, device=None):
    """The fp32 multipliers [rows, cols] descriptor `drop` applies (0 or 1/keep): what tests hand to the oracle."""
    ones = torch.ones((rows, cols), dtype=torch.float32, device=device or "cuda")
    out = torch.empty_like(ones)
    _lib.call("ofab_dropout_apply", _p(ones), _p(out), F32, rows, cols, _dp(drop), _s())
    return out


# ------------------------------------------------------------------------------------ reductions
def colsum(x2d, out_dtype=torch.bfloat16, out=None, accumulate=False, scratch=None):
    """out[c] (+)= sum_r x2d[r, c]; x2d may have a row stride (last dim contiguous)."""
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    rows, cols = x2d.shape
    if out is None:
        out = torch.empty(cols, dtype=out_dtype, device=x2d.device)
    if scratch is None:
        scratch = torch.empty(_lib.lib().ofab_colsum_scratch_elems(cols), dtype=torch.float32, device=x2d.device)
    _lib.call("ofab_colsum", _p(x2d), _DT[x2d.dtype], rows, cols, x2d.stride(0), _p(out), _DT[out.dtype], int(accumulate), _p(scratch), _s())
    return out


def _reduce_partials(partial, dtype):
    """[ns, rows, cols] fp32 partial slabs -> [ns, cols] `dtype` in one launch."""
    ns, _, cols = partial.shape
    out = torch.empty((ns, cols), dtype=dtype, device=partial.device)
    _lib.call("ofab_reduce_partials", _p(partial), ns, cols, _p(out), _DT[dtype], _s())
    return out


# Bias-gradient side channel: the LayerNorm backward kernels already hold every element of the gradient they
# emit (dx / da) in registers, so they also accumulate its column sums -- which is exactly the bias gradient of the
# Linear layer that consumes that gradient next in backward.  The hint is looked up by the gradient's address and is
# valid only for THE SAME tensor object (weak reference): the caching allocator hands a freed gradient's address to
# later tensors of the same size, and an unconsumed hint must never be taken for one of those.
_BIAS_HINTS = {}


def _hint_bias_grad(grad_tensor, colsum_vec):
    if len(_BIAS_HINTS) > 64:
        _BIAS_HINTS.clear()
    _BIAS_HINTS[(grad_tensor.data_ptr(), grad_tensor.numel())] = (weakref.ref(grad_tensor), colsum_vec, grad_tensor._version)


def _take_bias_hint(grad_tensor, n):
    """The hint is valid only for the same tensor object AND the same contents: autograd accumulates IN PLACE into the
    first-arrived gradient of a tensor with several consumers (and tensor hooks may edit a gradient), which keeps the
    object identity but bumps its version counter -- the column sums recorded before that are stale."""
    ent = _BIAS_HINTS.pop((grad_tensor.data_ptr(), grad_tensor.numel()), None)
    if ent is None or ent[0]() is not grad_tensor or ent[2] != grad_tensor._version:
        return None
    return ent[1] if ent[1].numel() == n else None


def cast_bf16(x):
    _need_cuda(x)
    x = _c(x)
    y = torch.empty_like(x, dtype=torch.bfloat16)
    _lib.call("ofab_cast_f32_bf16", _p(x), _p(y), x.numel(), _s())
    return y


def cast_f32(x):
    _need_cuda(x)
    x = _c(x)
    y = torch.empty_like(x, dtype=torch.float32)
    _lib.call("ofab_cast_bf16_f32", _p(x), _p(y), x.numel(), _s())
    return y


class _CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, to_f32):
        ctx.to_f32 = to_f32
        return cast_f32(x) if to_f32 else cast_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return (cast_bf16(g) if ctx.to_f32 else cast_f32(g)), None


def to_f32(x):
    return x if x.dtype == torch.float32 else _CastFn.apply(x, True)


def to_bf16(x):
    return x if x.dtype == torch.bfloat16 else _CastFn.apply(x, False)


# ------------------------------------------------------------------------------------ LayerNorm
def _partial_rows():
    return _lib.lib().ofab_ln_partial_rows()


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, gelu, out_dtype, drop=None):
        _need_cuda(x, weight, bias)
        x = _c(x)
        cols = x.shape[-1]
        rows = x.numel() // cols
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.call("ofab_ln_fwd", _p(x), _DT[x.dtype], _p(weight), _p(bias), _p(y), _DT[out_dtype], _p(mean), _p(rstd), rows, cols, eps, int(gelu), _dp(drop), _s())
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.gelu = gelu
        ctx.drop = drop
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        cols = x.shape[-1]
        rows = x.numel() // cols
        dx = torch.empty_like(x)
        partial = torch.empty((3, _partial_rows(), cols), dtype=torch.float32, device=x.device)
        _lib.call("ofab_ln_bwd", _p(dy), _DT[dy.dtype], _p(x), _DT[x.dtype], _p(weight), _p(mean), _p(rstd), _p(dx), _DT[dx.dtype], 0,
                  _p(partial), rows, cols, int(ctx.gelu), _dp(ctx.drop), _s())
        g = _reduce_partials(partial, weight.dtype)
        if dx.dtype == torch.bfloat16:
            _hint_bias_grad(dx, g[2])
        return dx, g[0], g[1], None, None, None, None


def layer_norm(x, weight, bias, eps=1e-5, gelu=False, out_dtype=torch.bfloat16, drop=None):
    """LN(drop?(gelu?(x))).  x fp32 -> bf16/fp32, or bf16 -> bf16/fp32 (gelu only bf16 -> bf16; `drop` = activation
    dropout descriptor, gelu form only)."""
    return _LayerNormFn.apply(x, weight, bias, eps, gelu, out_dtype, drop)


class _LnResLnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, w1, b1, w2, b2, eps, drop=None):
        _need_cuda(a, x)
        a, x = _c(a), _c(x)
        assert a.dtype == torch.bfloat16 and x.dtype == torch.float32
        cols = a.shape[-1]
        rows = a.numel() // cols
        x_new = torch.empty_like(x)
        y = torch.empty_like(a)
        stats = torch.empty((4, rows), dtype=torch.float32, device=a.device)
        _lib.call("ofab_ln_res_ln_fwd", _p(a), _p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(x_new), _p(y), _p(stats), rows, cols, eps, _dp(drop), _s())
        ctx.save_for_backward(a if w1 is not None else None, x_new, w1, w2, stats)
        ctx.shape_a = a.shape
        ctx.drop = drop
        return x_new, y

    @staticmethod
    def backward(ctx, dx_new, dy):
        a, x_new, w1, w2, stats = ctx.saved_tensors
        cols = x_new.shape[-1]
        rows = x_new.numel() // cols
        dx_new = torch.zeros_like(x_new) if dx_new is None else _c(dx_new)
        dy = torch.zeros(ctx.shape_a, dtype=torch.bfloat16, device=x_new.device) if dy is None else _c(dy)
        dx_tot = torch.empty_like(x_new)
        da = torch.empty(ctx.shape_a, dtype=torch.bfloat16, device=x_new.device)
        partial = torch.empty((5, _partial_rows(), cols), dtype=torch.float32, device=x_new.device)
        _lib.call("ofab_ln_res_ln_bwd", _p(dx_new), _p(dy), _p(a), _p(x_new), _p(w1), _p(w2), _p(stats), _p(dx_tot), _p(da), _p(partial), rows, cols, _dp(ctx.drop), _s())
        g = _reduce_partials(partial, w2.dtype)
        _hint_bias_grad(da, g[4])
        if w1 is None:
            return da, dx_tot, None, None, g[2], g[3], None, None
        return da, dx_tot, g[0], g[1], g[2], g[3], None, None


def ln_res_ln(a, x, w1, b1, w2, b2, eps=1e-5, drop=None):
    """x_new = x + drop(LN1(a)); y = LN2(x_new).  Returns (x_new fp32, y bf16).  w1 = b1 = None: x_new = x + drop(a).
    `drop`: residual dropout / drop-path descriptor of the block (DropoutState.spec) or None."""
    return _LnResLnFn.apply(a, x, w1, b1, w2, b2, eps, drop)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _c(x), _c(y)
        out = torch.empty_like(x)
        _lib.call("ofab_add_f32", _p(x), _p(y), _p(out), x.numel(), _s())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add_residual(x, y):
    """x (fp32) + y (bf16 or fp32) -> fp32: the un-fused residual add of the reference-layout API path."""
    return _AddFn.apply(x, to_f32(y))


# ------------------------------------------------------------------------------------ GEMM
def gemm(M, N, K, A, lda, a_mn, B, ldb, b_mn, out, ldd, bias=None, residual=None, ldr=0):
    """out[M,N] = A * B^T (+bias)(+residual) on tcgen05 (see include/ofab.h)."""
    _lib.call("ofab_gemm_bf16", M, N, K, _p(A), lda, int(a_mn), _p(B), ldb, int(b_mn), _p(bias), _p(residual), ldr, _p(out), ldd, _DT[out.dtype], _s())
    return out


def gemm_splitk(M, N, K, A, lda, a_mn, B, ldb, b_mn, out, ldd):
    """out[M,N] = A * B^T for few-tile / long-K problems (weight gradients): split-K partial slabs + one reduction
    when that pays, the plain kernel otherwise (see include/ofab.h)."""
    n_ws = _lib.lib().ofab_gemm_splitk_workspace_elems(M, N, K)
    ws = torch.empty(n_ws, dtype=torch.float32, device=out.device) if n_ws > 0 else None
    _lib.call("ofab_gemm_bf16_splitk", M, N, K, _p(A), lda, int(a_mn), _p(B), ldb, int(b_mn), _p(out), ldd, _DT[out.dtype], _p(ws), n_ws, _s())
    return out


def _grad_slot(w, shape):
    from .distributed import grad_slot

    return grad_slot(w, shape)


def _as2d(x):
    if x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 8 == 0:
        return x
    return _c(x).view(-1, x.shape[-1])


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        _need_cuda(x, weight)
        assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
        x2 = _as2d(x)
        M, K = x2.shape
        N = weight.shape[0]
        w = _c(weight)
        Np = (N + 7) // 8 * 8
        if residual is not None:
            r2 = _c(residual).view(M, N)
            buf = torch.empty((M, N), dtype=torch.float32, device=x.device)
            gemm(M, N, K, x2, x2.stride(0), 0, w, K, 0, buf, N, bias=bias, residual=r2, ldr=N)
            out = buf
        else:
            buf = torch.empty((M, Np), dtype=torch.bfloat16, device=x.device)
            # ragged N without bias (the tied vocabulary projection): compute the padded width -- rows of W beyond N
            # are zero-filled by TMA, so the pad columns hold zeros and the coalesced TMA-store epilogue applies
            gemm(M, Np if bias is None else N, K, x2, x2.stride(0), 0, w, K, 0, buf, Np, bias=bias)
            out = buf if Np == N else buf[:, :N]
        ctx.save_for_backward(x2, w)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.xshape = x.shape
        return out.view(*x.shape[:-1], N) if Np == N else out.unflatten(0, x.shape[:-1])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        d_res = dy if ctx.has_res else None
        dyb = dy if dy.dtype == torch.bfloat16 else cast_bf16(dy)
        dy2 = dyb.reshape(M, N) if dyb.is_contiguous() else _as2d(dyb.reshape(M, N))
        ld = dy2.stride(0)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
            gemm(M, K, N, dy2, ld, 0, w, K, 1, dx, K)  # dX = dY * W   (W read MN-major in place)
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = _grad_slot(w, (N, K))  # straight into the data-parallel gradient arena when one is active
            if dw is None:
                dw = torch.empty((N, K), dtype=torch.bfloat16, device=dy.device)
            gemm_splitk(N, K, M, dy2, ld, 1, x2, x2.stride(0), 1, dw, K)  # dW = dY^T * X  (both read transposed in place)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _take_bias_hint(dy, N)  # produced for free by the LayerNorm backward that emitted dy
            if db is None:
                # (measured: running this HBM-bound reduction on a side stream next to the two GEMMs above gains nothing --
                # the persistent GEMM CTAs leave no SM idle and the reduction's traffic slows them by as much)
                db = colsum(dy2, torch.bfloat16)
        return dx, dw, db, d_res


def linear(x, weight, bias=None, residual=None):
    """y = x W^T + b (bf16) ; with `residual` (fp32): y = residual + x W^T + b in fp32."""
    return _LinearFn.apply(x, weight, bias, residual)


class _PackedParamsFn(torch.autograd.Function):
    """Several parameters that sit back to back in ONE storage (pack_params) seen as a single [sum(rows), ...] operand, and
    their gradients as slices of ONE gradient tensor: the packed q|k|v (or k|v) projection runs as one GEMM without a
    per-forward torch.cat of the weights and without splitting copies of the packed gradient in backward."""

    @staticmethod
    def forward(ctx, *ps):
        first = ps[0]
        rows = sum(p.shape[0] for p in ps)
        out = first.detach().new_empty(0).set_(first.untyped_storage(), first.storage_offset(), (rows,) + tuple(first.shape[1:]))
        ctx.rows = [p.shape[0] for p in ps]
        return out

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        outs, o = [], 0
        for r in ctx.rows:
            outs.append(g[o:o + r])
            o += r
        return tuple(outs)


def pack_params(params):
    """Returns the parameters viewed as one tensor (rows concatenated).  The first call after the parameters were
    (re)allocated -- construction, .to(device / dtype) -- moves them into one shared storage (their .data become views of it);
    later calls are pointer arithmetic only.  Names, shapes and values of the parameters are unchanged."""
    first = params[0]
    es = first.element_size()
    adj = all(p.is_contiguous() and p.dtype == first.dtype and p.device == first.device for p in params)
    if adj:
        ptr = first.data_ptr()
        for p in params:
            if p.data_ptr() != ptr or p.untyped_storage().data_ptr() != first.untyped_storage().data_ptr():
                adj = False
                break
            ptr += p.numel() * es
    if not adj:
        with torch.no_grad():
            buf = torch.cat([p.detach().reshape(-1) for p in params])
            o = 0
            for p in params:
                p.data = buf[o:o + p.numel()].view(p.shape)
                o += p.numel()
    return _PackedParamsFn.apply(*params)


class _ScaleColsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, c, group):
        W, c = _c(W), _c(c)
        out = torch.empty_like(W)
        _lib.call("ofab_scale_cols", _p(W), _p(c), _p(out), W.shape[0], W.shape[1], group, _s())
        ctx.save_for_backward(W, c)
        ctx.group = group
        return out

    @staticmethod
    def backward(ctx, dWe):
        W, c = ctx.saved_tensors
        dWe = _c(dWe)
        dW = _grad_slot(W, W.shape)
        if dW is None:
            dW = torch.empty_like(W)
        dc = torch.zeros(c.shape, dtype=torch.float32, device=W.device)
        _lib.call("ofab_scale_cols_bwd", _p(dWe), _p(W), _p(c), _p(dW), _p(dc), W.shape[0], W.shape[1], ctx.group, _s())
        return dW, cast_bf16(dc), None


def scale_cols(W, c, group):
    """W_eff[n, k] = W[n, k] * c[k // group]: folds the per-head c_attn into out_proj."""
    return _ScaleColsFn.apply(W, c, group)


# ------------------------------------------------------------------------------------ attention
class PositionBias:
    """Structured replacement of the reference's dense attn_bias [B*H, T, S]:
    pq/pk : bf16 [1 or B, T, d] absolute-position projections (scale applied in-kernel), or None
    rp_idx: int32 [Tq, Tk] bucket ids (-1 = no relative bias), or None
    table : bf16 [n_buckets, H] this layer's relative-position table(s), or None
    """

    def __init__(self, pq=None, pk=None, rp_idx=None, table=None, abs=None):
        self.pq, self.pk, self.rp_idx, self.table = pq, pk, rp_idx, table
        # dense form of the abs-pos term: fp32 [H, Tq, ld] = pq_h . pk_h per head (abs_pos()), shared by every layer of a
        # forward.  When present the attention runs on the tcgen05 kernels with ONE additive fp16 tile per layer
        # (abs * scale + table[rp_idx]) instead of per-score gathers; when absent (e.g. the one-row slices of incremental
        # decoding) the structured mma.sync kernels are used.
        self.abs = abs


class _AbsPosFn(torch.autograd.Function):
    """abs[h, i, j] = pq[i, h*64:(h+1)*64] . pk[j, h*64:(h+1)*64]  (fp32 [H, Tq, ld], ld = ceil8(Tk); unscaled): the
    absolute-position term of OFAGeneralAdaptor.build_abs_pos_bias (adaptor/general.py:223-243) /
    TransformerDecoder.get_cross_pos_info (model/transformer.py:280-299) for ONE batch element -- it does not depend on b."""

    @staticmethod
    def forward(ctx, pq, pk, H):
        _need_cuda(pq, pk)
        pq, pk = _c(pq), _c(pk)
        assert pq.dtype == torch.bfloat16 and pq.shape[0] == 1 and pk.shape[0] == 1
        Tq, Tk, d = pq.shape[1], pk.shape[1], pq.shape[2]
        ld = (Tk + 7) // 8 * 8
        out = torch.zeros((H, Tq, ld), dtype=torch.float32, device=pq.device)
        q2, k2 = pq[0], pk[0]
        for h in range(H):
            gemm(Tq, Tk, 64, q2[:, h * 64:], d, 0, k2[:, h * 64:], d, 0, out[h], ld)
        ctx.save_for_backward(pq, pk)
        ctx.H = H
        return out

    @staticmethod
    def backward(ctx, dabs):
        pq, pk = ctx.saved_tensors
        H = ctx.H
        Tq, Tk, d = pq.shape[1], pk.shape[1], pq.shape[2]
        g = cast_bf16(_c(dabs))  # [H, Tq, ld]
        ld = g.shape[2]
        dpq = torch.empty_like(pq)
        dpk = torch.empty_like(pk)
        q2, k2 = pq[0], pk[0]
        for h in range(H):
            gemm(Tq, 64, Tk, g[h], ld, 0, k2[:, h * 64:], d, 1, dpq[0][:, h * 64:], d)   # dpq_h = dabs_h pk_h
            gemm(Tk, 64, Tq, g[h], ld, 1, q2[:, h * 64:], d, 1, dpk[0][:, h * 64:], d)   # dpk_h = dabs_h^T pq_h
        return dpq, dpk, None


def abs_pos(pq, pk, H):
    """Dense abs-pos term for PositionBias(abs=...): fp32 [H, Tq, ceil8(Tk)]."""
    return _AbsPosFn.apply(pq, pk, H)


_IDX16 = {}


def _idx16(rp_idx):
    """int32 [Tq, Tk] bucket map -> (int16 [Tq, ld], int16 [Tk, ld_t]): the packed pair layout the kernels gather from
    (row lengths padded to even with -1) and its transpose for the dK/dV kernel.  Cached per map (adaptors cache theirs)."""
    key = (rp_idx.data_ptr(), tuple(rp_idx.shape), rp_idx._version)
    ent = _IDX16.get(key)
    if ent is not None and ent[0]() is rp_idx:
        return ent[1], ent[2]
    if len(_IDX16) > 64:
        _IDX16.clear()
    Tq, Tk = rp_idx.shape
    a = torch.full((Tq, (Tk + 1) // 2 * 2), -1, dtype=torch.int16, device=rp_idx.device)
    a[:, :Tk] = rp_idx
    b = torch.full((Tk, (Tq + 1) // 2 * 2), -1, dtype=torch.int16, device=rp_idx.device)
    b[:, :Tq] = rp_idx.t()
    _IDX16[key] = (weakref.ref(rp_idx), a, b)
    return a, b


def _fill_attn(args, B, H, Tq, Tk, q, k, v, pq, pk, rp_idx, table_f32, kpm, causal, scale, o, lse, drop=None, bias=None, bias_t=None):
    args.B, args.H, args.Tq, args.Tk = B, H, Tq, Tk
    if bias is not None:
        args.bias, args.bias_hs, args.bias_ld = bias.data_ptr(), bias.stride(0), bias.shape[2]
        args.bias_t, args.bias_t_hs, args.bias_t_ld = bias_t.data_ptr(), bias_t.stride(0), bias_t.shape[2]
    else:
        args.bias = args.bias_t = None
        args.bias_hs = args.bias_t_hs = 0
        args.bias_ld = args.bias_t_ld = 0
    args.drop = None if drop is None else ctypes.pointer(drop)
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    args.q_bs, args.q_rs = q.stride(0), q.stride(1)
    args.k_bs, args.k_rs = k.stride(0), k.stride(1)
    args.v_bs, args.v_rs = v.stride(0), v.stride(1)
    if pq is not None:
        args.pq, args.pk = pq.data_ptr(), pk.data_ptr()
        args.pq_bs = 0 if pq.shape[0] == 1 else pq.stride(0)
        args.pq_rs = pq.stride(1)
        args.pk_bs = 0 if pk.shape[0] == 1 else pk.stride(0)
        args.pk_rs = pk.stride(1)
    else:
        args.pq = args.pk = None
    if rp_idx is not None:
        i16, i16t = _idx16(rp_idx)
        args._keep_idx = (i16, i16t)
        args.rp_idx, args.rp_ld, args.rp_idx_t, args.rp_ld_t = i16.data_ptr(), i16.shape[1], i16t.data_ptr(), i16t.shape[1]
        args.table, args.n_buckets = table_f32.data_ptr(), table_f32.shape[0]
    else:
        args.rp_idx = args.rp_idx_t = args.table = None
        args.n_buckets = args.rp_ld = args.rp_ld_t = 0
    args.kpm = None if kpm is None else kpm.data_ptr()
    args.causal = int(causal)
    args.scale = scale
    args.o, args.o_bs, args.o_rs = o.data_ptr(), o.stride(0), o.stride(1)
    args.lse = lse.data_ptr()


class _AttentionFn(torch.autograd.Function):
    """q_src: self-attention -> packed qkv [B, T, 3d]; cross-attention -> q [B, Tq, d] with kv_src [B, Tk, 2d]."""

    @staticmethod
    def forward(ctx, q_src, kv_src, pq, pk, table, rp_idx, kpm, causal, scale, H, drop=None, abs_t=None):
        _need_cuda(q_src)
        # the kernels read q / k / v through (batch, row) strides: a column slice of a packed projection or a prefix of a
        # K|V cache is used in place as long as rows are contiguous and 16-byte aligned
        strided_ok = lambda t: t.dim() == 3 and t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.data_ptr() % 16 == 0
        q_src = q_src if strided_ok(q_src) else _c(q_src)
        d = H * 64
        B = q_src.shape[0]
        if kv_src is None:
            assert q_src.shape[-1] == 3 * d
            q, k, v = q_src[..., :d], q_src[..., d:2 * d], q_src[..., 2 * d:]
        else:
            kv_src = kv_src if strided_ok(kv_src) else _c(kv_src)
            assert q_src.shape[-1] == d and kv_src.shape[-1] == 2 * d
            q, k, v = q_src, kv_src[..., :d], kv_src[..., d:]
        Tq, Tk = q.shape[1], k.shape[1]
        dense = abs_t is not None
        bias = bias_t = None
        if dense:
            # ONE additive tile per layer for all batch elements: scale * abs + table[rp_idx] (fp16) and its transpose
            assert abs_t.dtype == torch.float32 and abs_t.shape[0] == H and abs_t.shape[1] == Tq and abs_t.shape[2] >= Tk and abs_t.is_contiguous()
            ld, ld_t = (Tk + 7) // 8 * 8, (Tq + 7) // 8 * 8
            bias = torch.empty((H, Tq, ld), dtype=torch.float16, device=q_src.device)
            bias_t = torch.empty((H, Tk, ld_t), dtype=torch.float16, device=q_src.device)
            if rp_idx is not None:
                assert rp_idx.dtype == torch.int32 and rp_idx.shape == (Tq, Tk) and rp_idx.is_contiguous()
                table = _c(table)
            _lib.call("ofab_attn_bias_build", _p(abs_t), abs_t.shape[2], scale, _p(rp_idx), _p(table) if rp_idx is not None else None,
                      _DT[table.dtype] if rp_idx is not None else F32, 0 if rp_idx is None else table.shape[0], H, Tq, Tk,
                      _p(bias), ld, _p(bias_t), ld_t, _s())
            pq = pk = None
        elif pq is not None:
            pq, pk = _c(pq), _c(pk)
        table_f = None
        if rp_idx is not None and not dense:
            assert rp_idx.dtype == torch.int32 and rp_idx.shape == (Tq, Tk) and rp_idx.is_contiguous()
            table_f = _c(table) if table.dtype == torch.float32 else cast_f32(table)
        if kpm is not None:
            kpm = _c(kpm.to(torch.uint8) if kpm.dtype != torch.uint8 else kpm)
        o = torch.empty((B, Tq, d), dtype=torch.bfloat16, device=q_src.device)
        lse = torch.empty((B, H, Tq), dtype=torch.float32, device=q_src.device)
        a = _lib.AttnFwdArgs()
        _fill_attn(a, B, H, Tq, Tk, q, k, v, pq, pk, None if dense else rp_idx, table_f, kpm, causal, scale, o, lse, drop, bias, bias_t)
        _lib.call("ofab_attn_fwd", ctypes.byref(a), _s())
        ctx.save_for_backward(q_src, kv_src, pq, pk, table_f, rp_idx, kpm, o, lse, bias, bias_t)
        ctx.meta = (causal, scale, H, None if table is None else table.dtype, dense,
                    None if abs_t is None else tuple(abs_t.shape), None if (table is None or not dense) else tuple(table.shape))
        ctx.drop = drop
        return o

    @staticmethod
    def backward(ctx, d_o):
        q_src, kv_src, pq, pk, table_f, rp_idx, kpm, o, lse, bias, bias_t = ctx.saved_tensors
        causal, scale, H, table_dtype, dense, abs_shape, table_shape = ctx.meta
        d = H * 64
        B = q_src.shape[0]
        d_o = _c(d_o)
        if kv_src is None:
            q, k, v = q_src[..., :d], q_src[..., d:2 * d], q_src[..., 2 * d:]
            dq_src = torch.empty(q_src.shape, dtype=q_src.dtype, device=q_src.device)
            dq, dk, dv = dq_src[..., :d], dq_src[..., d:2 * d], dq_src[..., 2 * d:]
            dkv_src = None
        else:
            q, k, v = q_src, kv_src[..., :d], kv_src[..., d:]
            dq_src = torch.empty(q_src.shape, dtype=q_src.dtype, device=q_src.device)
            dkv_src = torch.empty(kv_src.shape, dtype=kv_src.dtype, device=kv_src.device)
            dq, dk, dv = dq_src, dkv_src[..., :d], dkv_src[..., d:]
        Tq, Tk = q.shape[1], k.shape[1]
        a = _lib.AttnBwdArgs()
        _fill_attn(a.f, B, H, Tq, Tk, q, k, v, pq, pk, None if dense else rp_idx, table_f, kpm, causal, scale, o, lse, ctx.drop, bias, bias_t)
        a.d_o, a.do_bs, a.do_rs = d_o.data_ptr(), d_o.stride(0), d_o.stride(1)
        ds = None
        if dense:  # dS per batch element; columns past a causal tile's last key block are never written
            ds = (torch.zeros if causal else torch.empty)((B, H, Tq, bias.shape[2]), dtype=torch.bfloat16, device=d_o.device)
        a.ds = None if ds is None else ds.data_ptr()
        a.dq, a.dk, a.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
        a.dq_bs, a.dq_rs = dq.stride(0), dq.stride(1)
        a.dk_bs, a.dk_rs = dk.stride(0), dk.stride(1)
        a.dv_bs, a.dv_rs = dv.stride(0), dv.stride(1)
        dpq = dpk = dtab = None
        if pq is not None:
            dpq = torch.empty((B, Tq, d), dtype=torch.bfloat16, device=d_o.device)
            dpk = torch.empty((B, Tk, d), dtype=torch.bfloat16, device=d_o.device)
            a.dpq, a.dpk = dpq.data_ptr(), dpk.data_ptr()
        else:
            a.dpq = a.dpk = None
        if rp_idx is not None and not dense:
            dtab = torch.zeros_like(table_f)
            a.dtable = dtab.data_ptr()
        else:
            a.dtable = None
        delta = torch.empty((B, H, Tq), dtype=torch.float32, device=d_o.device)
        a.delta = delta.data_ptr()
        # bias gradients of the q / k / v projections: the kernels leave one partial row of column sums per CTA (fp32, straight
        # from their accumulators) -- tiles of 128 rows on the tcgen05 path, 64 on the mma.sync one; the buffers are sized for
        # the finer tiling and zeroed, rows a path does not write add nothing (self-attention: Tq == Tk, one [3, B * tiles, d]
        # buffer whose slabs reduce straight into the packed [3d] vector)
        nq, nk = (Tq + 63) // 64, (Tk + 63) // 64
        if kv_src is None:
            part = torch.zeros((3, B * nq, d), dtype=torch.float32, device=d_o.device)
            a.dq_colsum, a.dk_colsum, a.dv_colsum = part[0].data_ptr(), part[1].data_ptr(), part[2].data_ptr()
        else:
            part_q = torch.zeros((1, B * nq, d), dtype=torch.float32, device=d_o.device)
            part_kv = torch.zeros((2, B * nk, d), dtype=torch.float32, device=d_o.device)
            a.dq_colsum, a.dk_colsum, a.dv_colsum = part_q.data_ptr(), part_kv[0].data_ptr(), part_kv[1].data_ptr()
        _lib.call("ofab_attn_bwd", ctypes.byref(a), _s())
        for grad, partial in ((dq_src, part) if kv_src is None else (dq_src, part_q), (None, None) if kv_src is None else (dkv_src, part_kv)):
            if grad is not None:
                vec = torch.empty(partial.shape[0] * d, dtype=torch.bfloat16, device=d_o.device)
                _lib.call("ofab_reduce_rows", _p(partial), partial.shape[0], partial.shape[1], d, _p(vec), BF16, _s())
                _hint_bias_grad(grad, vec)
        dabs = None
        if dense:  # ONE reduction over the batch per layer: table histogram + abs-pos gradient
            dabs = torch.zeros(abs_shape, dtype=torch.float32, device=d_o.device)
            if rp_idx is not None:
                dtab = torch.zeros(table_shape, dtype=torch.float32, device=d_o.device)
            _lib.call("ofab_attn_bias_bwd", _p(ds), B, H, Tq, Tk, ds.shape[3], _p(rp_idx), _p(dtab), _p(dabs), abs_shape[2], scale, 0, _s())
        if pq is not None:
            if pq.shape[0] == 1 and B > 1:  # broadcast positions: sum the per-sample grads
                dpq = colsum(dpq.view(B, Tq * d), torch.bfloat16).view(1, Tq, d)
            if pk.shape[0] == 1 and B > 1:
                dpk = colsum(dpk.view(B, Tk * d), torch.bfloat16).view(1, Tk, d)
        if dtab is not None:
            dtab = cast_bf16(dtab) if table_dtype == torch.bfloat16 else dtab
        return dq_src, dkv_src, dpq, dpk, dtab, None, None, None, None, None, None, dabs


def attention(q_src, kv_src, H, scale, bias: PositionBias = None, key_padding_mask=None, causal=False, drop=None):
    """`drop`: dropout descriptor for the attention probabilities (DropoutState.spec(p)) or None."""
    b = bias or PositionBias()
    if b.abs is not None:  # dense position tile: pq / pk enter through abs (their gradients flow through abs_pos())
        return _AttentionFn.apply(q_src, kv_src, None, None, b.table, b.rp_idx, key_padding_mask, causal, float(scale), H, drop, b.abs)
    return _AttentionFn.apply(q_src, kv_src, b.pq, b.pk, b.table, b.rp_idx, key_padding_mask, causal, float(scale), H, drop, None)


def attention_dropout_mask(drop, B, H, Tq, Tk, device=None):
    """The fp32 multipliers [B, H, Tq, Tk] an attention dropout descriptor applies to the probabilities (tests):
    recovered exactly by running the kernel on zero scores (uniform P = 1/Tk) against V = identity columns."""
    dev = device or "cuda"
    out = torch.empty((B, H, Tq, Tk), dtype=torch.float32, device=dev)
    q = torch.zeros((B, Tq, H * 64), dtype=torch.bfloat16, device=dev)
    for j0 in range(0, Tk, 64):
        n = min(64, Tk - j0)
        kv = torch.zeros((B, Tk, 2 * H * 64), dtype=torch.bfloat16, device=dev)
        v = kv[..., H * 64:].view(B, Tk, H, 64)
        for jj in range(n):
            v[:, j0 + jj, :, jj] = 1.0
        o = _AttentionFn.apply(q, kv, None, None, None, None, None, False, 1.0, H, drop, None)  # o[b,i,h,jj] = mask / Tk
        out[..., j0:j0 + n] = (o.view(B, Tq, H, 64)[..., :n].permute(0, 2, 1, 3) > 0).float() * (1.0 / (1.0 - drop.p))
    return out


# ------------------------------------------------------------------------------------ adaptor hook
class _EmbedLnFn(torch.autograd.Function):
    """out = LN(src + pos + type) (fp32 [B, T, d]); src = E[tokens] or [cls; dense]."""

    @staticmethod
    def forward(ctx, tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, eps, padding_idx, drop=None):
        dev = gamma.device
        _need_cuda(gamma)
        if tokens is not None:
            tokens = _c(tokens)
            B, T = tokens.shape
            d = E.shape[1]
            E = _c(E)
        else:
            dense = _c(dense)
            B, d = dense.shape[0], dense.shape[-1]
            T = dense.shape[1] + (1 if cls is not None else 0)
        if pos is not None:
            pos = _c(pos)
            assert pos.shape[0] >= T and pos.shape[-1] == d
        if zero_mask is not None:
            zero_mask = _c(zero_mask.to(torch.uint8))
        out = torch.empty((B, T, d), dtype=torch.float32, device=dev)
        mean = torch.empty(B * T, dtype=torch.float32, device=dev)
        rstd = torch.empty(B * T, dtype=torch.float32, device=dev)
        a = _lib.EmbedLnArgs()
        _EmbedLnFn._fill(a, B, T, d, tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, eps, out, mean, rstd, drop)
        _lib.call("ofab_embed_ln_fwd", ctypes.byref(a), _s())
        ctx.save_for_backward(tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, mean, rstd)
        ctx.meta = (B, T, d, eps, padding_idx)
        ctx.drop = drop
        return out

    @staticmethod
    def _fill(a, B, T, d, tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, eps, out, mean, rstd, drop=None):
        a.B, a.T, a.d = B, T, d
        a.drop = None if drop is None else ctypes.pointer(drop)
        a.tokens = None if tokens is None else tokens.data_ptr()
        a.E = None if E is None or tokens is None else E.data_ptr()
        a.dense = None if dense is None else dense.data_ptr()
        a.cls = None if cls is None else cls.data_ptr()
        a.has_cls = int(cls is not None)
        a.pos = None if pos is None else pos.data_ptr()
        a.type = None if type_vec is None else type_vec.data_ptr()
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
        a.zero_mask = None if zero_mask is None else zero_mask.data_ptr()
        a.eps = eps
        a.out = None if out is None else out.data_ptr()
        a.out_bs = T * d
        a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()

    @staticmethod
    def backward(ctx, dout):
        tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, mean, rstd = ctx.saved_tensors
        B, T, d, eps, padding_idx = ctx.meta
        dev = gamma.device
        dout = _c(dout)
        a = _lib.EmbedLnBwdArgs()
        _EmbedLnFn._fill(a.f, B, T, d, tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, eps, None, mean, rstd, ctx.drop)
        a.dout, a.dout_bs = dout.data_ptr(), T * d
        dE = ddense = dpos = None
        if tokens is not None and ctx.needs_input_grad[1]:
            dE = torch.zeros(E.shape, dtype=torch.float32, device=dev)
            a.dE = dE.data_ptr()
        else:
            a.dE = None
        a.padding_idx = -1 if padding_idx is None else padding_idx
        if dense is not None:
            ddense = torch.empty_like(dense)
            a.ddense = ddense.data_ptr()
        else:
            a.ddense = None
        if pos is not None and ctx.needs_input_grad[4]:
            dpos = torch.zeros((pos.shape[0], d), dtype=torch.float32, device=dev)
            a.dpos = dpos.data_ptr()
        else:
            a.dpos = None
        partial = torch.empty((4, _partial_rows(), d), dtype=torch.float32, device=dev)
        a.dgb_partial = partial.data_ptr()
        _lib.call("ofab_embed_ln_bwd", ctypes.byref(a), _s())
        g = _reduce_partials(partial, gamma.dtype)
        dgamma, dbeta = g[0], g[1]
        dtype_vec = g[2].view(type_vec.shape) if type_vec is not None else None
        dcls = g[3].view(cls.shape) if cls is not None else None
        if dE is not None:
            dE = cast_bf16(dE)
        if dpos is not None:
            dpos = cast_bf16(dpos).view(pos.shape)
        return None, dE, ddense, dcls, dpos, dtype_vec, dgamma, dbeta, None, None, None, None


def embed_ln(gamma, beta, tokens=None, E=None, dense=None, cls=None, pos=None, type_vec=None, zero_mask=None, eps=1e-5, padding_idx=None, drop=None):
    return _EmbedLnFn.apply(tokens, E, dense, cls, pos, type_vec, gamma, beta, zero_mask, eps, padding_idx, drop)


# ------------------------------------------------------------------------------------ criterion
class _CrossEntropyFn(torch.autograd.Function):
    """sum-reduced CE over rows whose target != ignore_index; logits bf16 [..., V] (row stride % 8 == 0)."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index, label_smoothing=0.0):
        _need_cuda(logits, target)
        V = logits.shape[-1]
        l2 = logits.reshape(-1, V) if logits.dim() != 2 else logits
        if l2.stride(1) != 1 or l2.stride(0) % 8 != 0:
            Vp = (V + 7) // 8 * 8
            buf = torch.empty((l2.shape[0], Vp), dtype=torch.bfloat16, device=logits.device)
            buf[:, :V].copy_(l2)
            l2 = buf[:, :V]
        rows, ld = l2.shape[0], l2.stride(0)
        tgt = _c(target.reshape(-1))
        lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        _lib.call("ofab_ce_fwd", _p(l2), rows, V, ld, _p(tgt), ignore_index, _p(lse), _p(loss), float(label_smoothing), None, _s())
        ctx.save_for_backward(l2, tgt, lse)
        ctx.meta = (ignore_index, logits.shape, float(label_smoothing))
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        l2, tgt, lse = ctx.saved_tensors
        ignore_index, shape, eps = ctx.meta
        rows, V = l2.shape
        ld = l2.stride(0)
        dl = torch.empty((rows, ld), dtype=torch.bfloat16, device=l2.device)
        gs = _c(g.reshape(1).to(torch.float32))
        _lib.call("ofab_ce_bwd", _p(l2), rows, V, ld, _p(tgt), ignore_index, _p(lse), _p(gs), _p(dl), eps, _s())
        return dl[:, :V].view(shape) if ld == V else dl[:, :V].unflatten(0, shape[:-1]), None, None, None


def cross_entropy_sum(logits, target, ignore_index=1, label_smoothing=0.0):
    """sum-reduced (label-smoothed) cross entropy over the rows whose target != ignore_index."""
    return _CrossEntropyFn.apply(logits, target, ignore_index, label_smoothing)


class _CrossEntropyRowsFn(torch.autograd.Function):
    """Per-row (label-smoothed) cross entropy with constraint masks: returns (row_loss, row_nll) fp32 [rows]."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index, label_smoothing, cmask, c_lo, c_hi):
        _need_cuda(logits, target)
        V = logits.shape[-1]
        l2 = logits.reshape(-1, V)
        if l2.dtype != torch.bfloat16 or l2.stride(1) != 1 or l2.stride(0) % 8 != 0:
            Vp = (V + 7) // 8 * 8
            buf = torch.zeros((l2.shape[0], Vp), dtype=torch.bfloat16, device=logits.device)
            buf[:, :V].copy_(l2)
            l2 = buf[:, :V]
        rows, ld = l2.shape[0], l2.stride(0)
        tgt = _c(target.reshape(-1))
        if cmask is not None:
            cmask = _c(cmask.reshape(rows, V).to(torch.uint8))
        dev = logits.device
        lse, nal, rl, rn = (torch.empty(rows, dtype=torch.float32, device=dev) for _ in range(4))
        a = _lib.CeRowsArgs()
        a.logits, a.rows, a.V, a.ld, a.target, a.ignore_index = l2.data_ptr(), rows, V, ld, tgt.data_ptr(), ignore_index
        a.cmask = None if cmask is None else cmask.data_ptr()
        a.c_lo, a.c_hi, a.label_smoothing = c_lo, c_hi, float(label_smoothing)
        a.lse, a.n_allowed, a.row_loss, a.row_nll = lse.data_ptr(), nal.data_ptr(), rl.data_ptr(), rn.data_ptr()
        _lib.call("ofab_ce_rows_fwd", ctypes.byref(a), _s())
        ctx.save_for_backward(l2, tgt, cmask, lse, nal)
        ctx.meta = (ignore_index, logits.shape, float(label_smoothing), c_lo, c_hi)
        ctx.mark_non_differentiable(rn)
        return rl, rn

    @staticmethod
    def backward(ctx, g_loss, _g_nll):
        l2, tgt, cmask, lse, nal = ctx.saved_tensors
        ignore_index, shape, eps, c_lo, c_hi = ctx.meta
        rows, V = l2.shape
        ld = l2.stride(0)
        dl = torch.empty((rows, ld), dtype=torch.bfloat16, device=l2.device)
        gs = _c(g_loss.to(torch.float32))
        a = _lib.CeRowsArgs()
        a.logits, a.rows, a.V, a.ld, a.target, a.ignore_index = l2.data_ptr(), rows, V, ld, tgt.data_ptr(), ignore_index
        a.cmask = None if cmask is None else cmask.data_ptr()
        a.c_lo, a.c_hi, a.label_smoothing = c_lo, c_hi, eps
        a.lse, a.n_allowed, a.row_scale, a.dlogits = lse.data_ptr(), nal.data_ptr(), gs.data_ptr(), dl.data_ptr()
        _lib.call("ofab_ce_rows_bwd", ctypes.byref(a), _s())
        d = dl[:, :V]
        return (d.view(shape) if ld == V else d.unflatten(0, shape[:-1])), None, None, None, None, None, None


def cross_entropy_rows(logits, target, ignore_index=1, label_smoothing=0.0, constraint_masks=None, constraint_range=None):
    """Per-row criterion of label_smoothed_cross_entropy.py:62-92 with constraint masks (bool [.., V], True = allowed) and / or
    `constraint_range` = (start, end): returns (row_loss, row_nll) fp32 [rows]; ignored rows give 0."""
    c_lo, c_hi = (-1, -1) if constraint_range is None else (int(constraint_range[0]), int(constraint_range[1]))
    return _CrossEntropyRowsFn.apply(logits, target, ignore_index, label_smoothing, constraint_masks, c_lo, c_hi)


class _LinearCrossEntropyFn(torch.autograd.Function):
    """loss = sum-CE(x E^T, target): the tied output projection (adaptor/base.py:131) fused with the
    criterion (cross_entropy.py:62-67) so the [rows, V] logits live only as one bf16 scratch."""

    @staticmethod
    def forward(ctx, x, E, target, ignore_index, label_smoothing=0.0, nll_out=None):
        _need_cuda(x, E, target)
        x2 = _as2d(x)
        M, K = x2.shape
        V = E.shape[0]
        Vp = (V + 7) // 8 * 8
        E = _c(E)
        logits = torch.empty((M, Vp), dtype=torch.bfloat16, device=x.device)
        gemm(M, Vp, K, x2, x2.stride(0), 0, E, K, 0, logits, Vp)  # pad columns [V, Vp) = 0 (TMA zero-fills rows of E beyond V)
        tgt = _c(target.reshape(-1))
        lse = torch.empty(M, dtype=torch.float32, device=x.device)
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        _lib.call("ofab_ce_fwd", _p(logits), M, V, Vp, _p(tgt), ignore_index, _p(lse), _p(loss), float(label_smoothing), _p(nll_out), _s())
        ctx.save_for_backward(x2, E, tgt, lse, logits)
        ctx.meta = (ignore_index, x.shape, float(label_smoothing))
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        x2, E, tgt, lse, logits = ctx.saved_tensors
        ignore_index, xshape, eps = ctx.meta
        if getattr(ctx, "consumed", False):
            raise RuntimeError("linear_cross_entropy: backward ran twice over the same graph (retain_graph): the logits scratch "
                               "was overwritten by its gradient in the first pass; call forward again")
        ctx.consumed = True
        M, K = x2.shape
        V = E.shape[0]
        Vp = logits.shape[1]
        gs = _c(g.reshape(1).to(torch.float32))
        _lib.call("ofab_ce_bwd", _p(logits), M, V, Vp, _p(tgt), ignore_index, _p(lse), _p(gs), _p(logits), eps, _s())  # in place
        dx = dE = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.bfloat16, device=x2.device)
            gemm(M, K, V, logits, Vp, 0, E, K, 1, dx, K)
            dx = dx.view(xshape)
        if ctx.needs_input_grad[1]:
            dE = torch.empty((V, K), dtype=torch.bfloat16, device=x2.device)
            gemm(V, K, M, logits, Vp, 1, x2, x2.stride(0), 1, dE, K)
        return dx, dE, None, None, None, None


class _ChunkedLinearCrossEntropyFn(torch.autograd.Function):
    """Same value and gradients as _LinearCrossEntropyFn without a [rows, V] scratch: the rows go through the projection and
    the criterion `chunk_rows` at a time, so only a [chunk_rows, V] bf16 tile ever exists (it stays in the 126 MB L2 between
    the kernels that touch it for chunk_rows <= 1024 at V ~ 59 k).  Forward keeps the per-row log-sum-exp only; backward
    recomputes each chunk's logits (one extra projection GEMM, the usual price of not storing them), turns them into their
    gradient in place and feeds the dX / dE GEMMs.  dE accumulates over the chunks in fp32 (the GEMM's residual input) and is
    rounded to bf16 once, like the one-shot form."""

    @staticmethod
    def forward(ctx, x, E, target, ignore_index, label_smoothing, chunk_rows, nll_out):
        _need_cuda(x, E, target)
        x2 = _as2d(x)
        M, K = x2.shape
        V = E.shape[0]
        Vp = (V + 7) // 8 * 8
        E = _c(E)
        dev = x.device
        R = max(8, min(int(chunk_rows), M))
        tile = torch.empty((R, Vp), dtype=torch.bfloat16, device=dev)
        tgt = _c(target.reshape(-1))
        lse = torch.empty(M, dtype=torch.float32, device=dev)
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        eps = float(label_smoothing)
        for r0 in range(0, M, R):
            m = min(R, M - r0)
            gemm(m, Vp, K, x2[r0:r0 + m], x2.stride(0), 0, E, K, 0, tile, Vp)  # pad columns [V, Vp) = 0 (TMA zero-fills rows of E beyond V)
            _lib.call("ofab_ce_fwd", _p(tile), m, V, Vp, _p(tgt[r0:r0 + m]), ignore_index, _p(lse[r0:r0 + m]), _p(loss), eps, _p(nll_out), _s())  # sums accumulate
        ctx.save_for_backward(x2, E, tgt, lse)
        ctx.meta = (ignore_index, x.shape, eps, R)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        x2, E, tgt, lse = ctx.saved_tensors
        ignore_index, xshape, eps, R = ctx.meta
        M, K = x2.shape
        V = E.shape[0]
        Vp = (V + 7) // 8 * 8
        dev = x2.device
        need_dx, need_dE = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gs = _c(g.reshape(1).to(torch.float32))
        tile = torch.empty((R, Vp), dtype=torch.bfloat16, device=dev)
        dx = torch.empty((M, K), dtype=torch.bfloat16, device=dev) if need_dx else None
        dE32 = torch.empty((V, K), dtype=torch.float32, device=dev) if need_dE else None
        for r0 in range(0, M, R):
            m = min(R, M - r0)
            xs, ts = x2[r0:r0 + m], tgt[r0:r0 + m]
            gemm(m, Vp, K, xs, x2.stride(0), 0, E, K, 0, tile, Vp)
            _lib.call("ofab_ce_bwd", _p(tile), m, V, Vp, _p(ts), ignore_index, _p(lse[r0:r0 + m]), _p(gs), _p(tile), eps, _s())  # in place
            if need_dx:
                gemm_splitk(m, K, V, tile, Vp, 0, E, K, 1, dx[r0:r0 + m], K)  # few output tiles, contraction over the vocabulary
            if need_dE:
                gemm(V, K, m, tile, Vp, 1, xs, x2.stride(0), 1, dE32, K, residual=dE32 if r0 > 0 else None, ldr=K)
        return (None if dx is None else dx.view(xshape)), (cast_bf16(dE32) if need_dE else None), None, None, None, None, None


def linear_cross_entropy(x, E, target, ignore_index=1, label_smoothing=0.0, nll_out=None, chunk_rows=None):
    """loss = sum-CE(x E^T, target), optionally label-smoothed (label_smoothed_cross_entropy.py:62-92).  nll_out: optional
    zeroed fp32 [1] tensor that receives the plain nll sum (the reference logs it next to the loss).
    chunk_rows (or env OFAB_CE_CHUNK_ROWS): process the rows in chunks of that many so that no [rows, V] scratch is allocated
    (412 MB at B=64, T=64, V=50265); default: one shot."""
    if chunk_rows is None and os.environ.get("OFAB_CE_CHUNK_ROWS"):
        chunk_rows = int(os.environ["OFAB_CE_CHUNK_ROWS"])
    if chunk_rows:
        return _ChunkedLinearCrossEntropyFn.apply(x, E, target, ignore_index, label_smoothing, int(chunk_rows), nll_out)
    return _LinearCrossEntropyFn.apply(x, E, target, ignore_index, label_smoothing, nll_out)


class _CtcFn(torch.autograd.Function):
    """sum over utterances of the CTC negative log likelihood of `logits` (fused log-softmax)."""

    @staticmethod
    def forward(ctx, logits, input_lengths, targets, target_lengths, blank, zero_infinity, time_major):
        _need_cuda(logits, targets, target_lengths)
        assert logits.dim() == 3 and logits.stride(2) == 1 and logits.dtype in _DT
        T, B = (logits.shape[0], logits.shape[1]) if time_major else (logits.shape[1], logits.shape[0])
        ts, bs = (logits.stride(0), logits.stride(1)) if time_major else (logits.stride(1), logits.stride(0))
        C = logits.shape[2]
        targets = _c(targets.to(torch.int64))
        Lmax = targets.shape[1]
        target_lengths = _c(target_lengths.to(device=logits.device, dtype=torch.int64))
        if input_lengths is not None:
            input_lengths = _c(input_lengths.to(device=logits.device, dtype=torch.int64))
        dev = logits.device
        lse = torch.empty((B, T), dtype=torch.float32, device=dev)
        alpha = torch.empty((B, T, 2 * Lmax + 1), dtype=torch.float32, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        a = _lib.CtcArgs()
        a.logits, a.dt, a.t_stride, a.b_stride = logits.data_ptr(), _DT[logits.dtype], ts, bs
        a.B, a.T, a.C = B, T, C
        a.input_lengths = None if input_lengths is None else input_lengths.data_ptr()
        a.targets, a.Lmax, a.target_lengths = targets.data_ptr(), Lmax, target_lengths.data_ptr()
        a.blank, a.zero_infinity = int(blank), int(zero_infinity)
        a.lse, a.alpha, a.nll = lse.data_ptr(), alpha.data_ptr(), nll.data_ptr()
        _lib.call("ofab_ctc_fwd", ctypes.byref(a), _s())
        ctx.save_for_backward(logits, input_lengths, targets, target_lengths, lse, alpha, nll)
        ctx.args = a
        ctx.mark_non_differentiable(nll)
        per = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll) if zero_infinity else nll
        return per.sum(), nll

    @staticmethod
    def backward(ctx, g, _g_nll):
        logits, input_lengths, targets, target_lengths, lse, alpha, nll = ctx.saved_tensors
        a = ctx.args  # pointers of the saved tensors are unchanged
        gs = _c(g.reshape(1).to(torch.float32))
        dl = torch.empty_strided(logits.shape, logits.stride(), dtype=logits.dtype, device=logits.device)
        a.gscale, a.dlogits = gs.data_ptr(), dl.data_ptr()
        _lib.call("ofab_ctc_bwd", ctypes.byref(a), _s())
        return dl, None, None, None, None, None, None


def ctc_loss_sum(logits, targets, target_lengths, input_lengths=None, blank=0, zero_infinity=True, time_major=False):
    """CTC term of the ASR criterion (speech_to_text_loss.py:339-379) on raw logits [B, T, C] (time_major: [T, B, C]);
    targets int64 [B, Lmax] left-aligned.  Returns (loss_sum, nll_per_utterance)."""
    return _CtcFn.apply(logits, input_lengths, targets, target_lengths, blank, zero_infinity, time_major)


# ------------------------------------------------------------------------------------ adaptors' convs
def patch_im2col(img, patch, ldk):
    """[B, C, H, W] (fp32/bf16) -> bf16 [B*(H/p)*(W/p), ldk]; no gradient (the image is data)."""
    _need_cuda(img)
    img = _c(img)
    B, C, H, W = img.shape
    cols = torch.empty((B * (H // patch) * (W // patch), ldk), dtype=torch.bfloat16, device=img.device)
    _lib.call("ofab_patch_im2col", _p(img), _DT[img.dtype], B, C, H, W, patch, _p(cols), ldk, _s())
    return cols


class _TransposeLast2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w3):
        w3 = _c(w3)
        O, A, Bd = w3.shape
        out = torch.empty((O, Bd, A), dtype=w3.dtype, device=w3.device)
        _lib.call("ofab_transpose_last2", _p(w3), _p(out), O, A, Bd, _s())
        return out

    @staticmethod
    def backward(ctx, g):
        g = _c(g)
        O, Bd, A = g.shape
        out = torch.empty((O, A, Bd), dtype=g.dtype, device=g.device)
        _lib.call("ofab_transpose_last2", _p(g), _p(out), O, Bd, A, _s())
        return out


def transpose_last2(w3):
    return _TransposeLast2Fn.apply(w3)


class _Conv1ReluFn(torch.autograd.Function):
    """Conv2d(1, C, 3, stride 2) + ReLU on fbank [B, L, F] -> bf16 [B, H1, W1, C] (channel-last)."""

    @staticmethod
    def forward(ctx, fbank, w, b):
        _need_cuda(fbank, w)
        fbank = _c(fbank)
        B, L, F = fbank.shape
        C = w.shape[0]
        w9 = _c(w).view(C, 9)
        H1, W1 = (L - 3) // 2 + 1, (F - 3) // 2 + 1
        out = torch.empty((B, H1, W1, C), dtype=torch.bfloat16, device=fbank.device)
        _lib.call("ofab_conv1_relu_fwd", _p(fbank), _DT[fbank.dtype], B, L, F, _p(w9), _p(b), C, _p(out), _s())
        ctx.save_for_backward(fbank, out)
        ctx.wshape = w.shape
        return out

    @staticmethod
    def backward(ctx, dy):
        fbank, y = ctx.saved_tensors
        dy = _c(dy)
        B, L, F = fbank.shape
        C = y.shape[-1]
        dw = torch.zeros((C, 9), dtype=torch.float32, device=dy.device)
        db = torch.zeros(C, dtype=torch.float32, device=dy.device)
        _lib.call("ofab_conv1_relu_bwd", _p(fbank), _DT[fbank.dtype], B, L, F, _p(y), _p(dy), C, _p(dw), _p(db), _s())
        return None, cast_bf16(dw).view(ctx.wshape), cast_bf16(db)


def conv1_relu(fbank, w, b):
    return _Conv1ReluFn.apply(fbank, w, b)


class _Im2col3x3s2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        B, Hin, Win, C = x.shape
        Ho, Wo = (Hin - 3) // 2 + 1, (Win - 3) // 2 + 1
        cols = torch.empty((B * Ho * Wo, 9 * C), dtype=torch.bfloat16, device=x.device)
        _lib.call("ofab_im2col_3x3s2", _p(x), B, Hin, Win, C, _p(cols), _s())
        ctx.shape = x.shape
        return cols

    @staticmethod
    def backward(ctx, dcols):
        dcols = _c(dcols)
        B, Hin, Win, C = ctx.shape
        dx = torch.empty(ctx.shape, dtype=torch.bfloat16, device=dcols.device)
        _lib.call("ofab_col2im_3x3s2", _p(dcols), B, Hin, Win, C, _p(dx), _s())
        return dx


def im2col_3x3s2(x):
    return _Im2col3x3s2Fn.apply(x)


class _ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = _c(x).clone() if x.requires_grad or True else x
        _lib.call("ofab_relu_inplace", _p(y), y.numel(), _s())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy).clone()
        _lib.call("ofab_relu_bwd_inplace", _p(y), _p(dy), dy.numel(), _s())
        return dy


def relu(x):
    return _ReluFn.apply(x)


# ------------------------------------------------------------------------------------ ResNet pieces (NHWC bf16)
def video_frames(video):
    """clip [B, C, F, H, W] -> (frames bf16 [B*F, C, H, W], all-zero-frame mask bool [B, F]); data, no gradient."""
    _need_cuda(video)
    video = _c(video)
    B, C, F, H, W = video.shape
    frames = torch.empty((B * F, C, H, W), dtype=torch.bfloat16, device=video.device)
    zero = torch.empty((B, F), dtype=torch.uint8, device=video.device)
    _lib.call("ofab_video_frames", _p(video), _DT[video.dtype], B, C, F, H * W, _p(frames), _p(zero), _s())
    return frames, zero.view(torch.bool)


def im2col_nchw(img, k, stride, pad, ldk):
    """stem im2col from the NCHW image (data: no gradient)."""
    _need_cuda(img)
    img = _c(img)
    B, C, H, W = img.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    cols = torch.empty((B * Ho * Wo, ldk), dtype=torch.bfloat16, device=img.device)
    _lib.call("ofab_im2col_nchw", _p(img), _DT[img.dtype], B, C, H, W, k, stride, pad, _p(cols), ldk, _s())
    return cols, Ho, Wo


class _Im2colNhwcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        x = _c(x)
        B, H, W, C = x.shape
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        cols = torch.empty((B * Ho * Wo, k * k * C), dtype=torch.bfloat16, device=x.device)
        _lib.call("ofab_im2col_nhwc", _p(x), B, H, W, C, k, stride, pad, _p(cols), _s())
        ctx.meta = (x.shape, k, stride, pad)
        return cols

    @staticmethod
    def backward(ctx, dcols):
        shape, k, stride, pad = ctx.meta
        B, H, W, C = shape
        dcols = _c(dcols)
        dx = torch.empty(shape, dtype=torch.bfloat16, device=dcols.device)
        _lib.call("ofab_col2im_nhwc", _p(dcols), B, H, W, C, k, stride, pad, _p(dx), _s())
        return dx, None, None, None


def im2col_nhwc(x, k, stride, pad):
    return _Im2colNhwcFn.apply(x, k, stride, pad)


class _Subsample2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        B, H, W, C = x.shape
        y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=x.dtype, device=x.device)
        _lib.call("ofab_subsample2", _p(x), B, H, W, C, _p(y), 0, _s())
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W, C = ctx.shape
        dy = _c(dy)
        dx = torch.empty(ctx.shape, dtype=dy.dtype, device=dy.device)
        _lib.call("ofab_subsample2", _p(dy), B, H, W, C, _p(dx), 1, _s())
        return dx


def subsample2(x):
    return _Subsample2Fn.apply(x)


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        B, H, W, C = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, Ho, Wo, C), dtype=x.dtype, device=x.device)
        arg = torch.empty((B, Ho, Wo, C), dtype=torch.uint8, device=x.device)
        _lib.call("ofab_maxpool3x3s2_fwd", _p(x), B, H, W, C, _p(y), _p(arg), _s())
        ctx.save_for_backward(arg)
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        B, H, W, C = ctx.shape
        dy = _c(dy)
        dx = torch.empty(ctx.shape, dtype=dy.dtype, device=dy.device)
        _lib.call("ofab_maxpool3x3s2_bwd", _p(dy), _p(arg), B, H, W, C, _p(dx), _s())
        return dx


def maxpool3x3s2(x):
    return _MaxPoolFn.apply(x)


class _BatchNormFn(torch.autograd.Function):
    """Training-mode BatchNorm over the rows of x [.., C] fused with (+residual) and ReLU."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, relu, eps, momentum, run_mean, run_var):
        x = _c(x)
        C = x.shape[-1]
        R = x.numel() // C
        dev = x.device
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        var = torch.empty(C, dtype=torch.float32, device=dev)
        scratch = torch.empty(_lib.lib().ofab_bn_scratch_elems(C), dtype=torch.float32, device=dev)
        rdt = _DT[run_mean.dtype] if run_mean is not None else F32
        _lib.call("ofab_bn_stats", _p(x), R, C, _p(mean), _p(var), _p(run_mean), _p(run_var), rdt, momentum, _p(scratch), _s())
        res = None if residual is None else _c(residual)
        y = torch.empty_like(x)
        _lib.call("ofab_bn_apply", _p(x), _p(mean), _p(var), _p(gamma), _p(beta), _p(res), _p(y), R, C, eps, int(relu), _s())
        ctx.save_for_backward(x, y if relu else None, gamma, mean, var)
        ctx.meta = (relu, eps, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, var = ctx.saved_tensors
        relu, eps, has_res = ctx.meta
        dy = _c(dy)
        C = x.shape[-1]
        R = x.numel() // C
        dev = x.device
        sums = torch.empty((2, C), dtype=torch.float32, device=dev)
        scratch = torch.empty(_lib.lib().ofab_bn_scratch_elems(C), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        _lib.call("ofab_bn_bwd", _p(dy), _p(x), _p(y), _p(mean), _p(var), _p(gamma), _p(sums), _p(dx), _p(dres), R, C, eps, int(relu), _p(scratch), _s())
        g = cast_bf16(sums) if gamma.dtype == torch.bfloat16 else sums
        return dx, g[1], g[0], dres, None, None, None, None, None


def batch_norm_train(x, gamma, beta, residual=None, relu=False, eps=1e-5, momentum=0.1, run_mean=None, run_var=None):
    return _BatchNormFn.apply(x, gamma, beta, residual, relu, eps, momentum, run_mean, run_var)


class _BatchNormEvalFn(torch.autograd.Function):
    """Eval-mode BatchNorm (running statistics are constants) fused with (+residual) and ReLU."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, relu, eps, run_mean, run_var):
        x = _c(x)
        C = x.shape[-1]
        R = x.numel() // C
        mean = _c(run_mean) if run_mean.dtype == torch.float32 else cast_f32(run_mean)
        var = _c(run_var) if run_var.dtype == torch.float32 else cast_f32(run_var)
        res = None if residual is None else _c(residual)
        y = torch.empty_like(x)
        _lib.call("ofab_bn_apply", _p(x), _p(mean), _p(var), _p(gamma), _p(beta), _p(res), _p(y), R, C, eps, int(relu), _s())
        ctx.save_for_backward(x, y if relu else None, gamma, mean, var)
        ctx.meta = (relu, eps, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, var = ctx.saved_tensors
        relu, eps, has_res = ctx.meta
        dy = _c(dy)
        C = x.shape[-1]
        R = x.numel() // C
        dev = x.device
        want_affine = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        sums = torch.empty((2, C), dtype=torch.float32, device=dev) if want_affine else None
        scratch = torch.empty(_lib.lib().ofab_bn_scratch_elems(C), dtype=torch.float32, device=dev) if want_affine else None
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        _lib.call("ofab_bn_bwd_eval", _p(dy), _p(x), _p(y), _p(mean), _p(var), _p(gamma), _p(sums), _p(dx), _p(dres), R, C, eps, int(relu), _p(scratch), _s())
        dg = db = None
        if want_affine:
            g = cast_bf16(sums) if gamma.dtype == torch.bfloat16 else sums
            dg, db = g[1], g[0]
        return dx, dg, db, dres, None, None, None, None


def batch_norm_eval(x, gamma, beta, run_mean, run_var, residual=None, relu=False, eps=1e-5):
    """nn.BatchNorm2d in eval mode (model.eval() / freeze_resnet, adaptor/image_resnet.py:107-114) on channel-last x."""
    return _BatchNormEvalFn.apply(x, gamma, beta, residual, relu, eps, run_mean, run_var)
