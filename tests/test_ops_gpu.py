"""GPU: every kernel family of libofab against a plain PyTorch fp32 restatement of the same op on
the same (bf16-rounded) inputs.  Tolerances: fp32 outputs 2e-5 rel-L2 (accumulation order); bf16
outputs 6e-3 (one bf16 rounding, eps = 7.8e-3 per element)."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel_l2

pytestmark = pytest.mark.gpu

TOL32 = 2e-5
TOL16 = 6e-3


def dev():
    return torch.device("cuda:0")


def g():
    return torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16, scale=1.0, gen=None):
    return (torch.randn(*shape, generator=gen or g()) * scale).to(dtype).to(dev())


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from ofasys_b200 import _lib

    _lib.check(_lib.lib().ofab_device_check(0), "device check")


# ----------------------------------------------------------------------------------------- LN
@pytest.mark.parametrize("cols", [64, 128, 256, 768, 1024, 2048, 3072, 4096])
@pytest.mark.parametrize("rows", [1, 37, 1031])
def test_layer_norm_fwd_bwd(rows, cols):
    from ofasys_b200 import ops

    gen = g()
    x = rnd(rows, cols, dtype=torch.float32, gen=gen).requires_grad_(True)
    w = (torch.rand(cols, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    b = (torch.randn(cols, generator=gen) * 0.1).bfloat16().to(dev()).requires_grad_(True)
    dy = rnd(rows, cols, gen=gen)
    y = ops.layer_norm(x, w, b)
    y.backward(dy)
    xr = x.detach().clone().requires_grad_(True)
    wr, br = w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    yr = F.layer_norm(xr, (cols,), wr, br, 1e-5)
    yr.backward(dy.float())
    assert rel_l2(y, yr) <= TOL16
    assert rel_l2(x.grad, xr.grad) <= 1e-4
    assert rel_l2(w.grad, wr.grad) <= TOL16 and rel_l2(b.grad, br.grad) <= TOL16


@pytest.mark.parametrize("cols", [512, 1024, 3072, 4096])
def test_gelu_layer_norm(cols):
    from ofasys_b200 import ops

    gen = g()
    rows = 517
    h = rnd(rows, cols, gen=gen).requires_grad_(True)
    w = (torch.rand(cols, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    b = (torch.randn(cols, generator=gen) * 0.1).bfloat16().to(dev()).requires_grad_(True)
    dy = rnd(rows, cols, gen=gen)
    y = ops.layer_norm(h, w, b, gelu=True)
    y.backward(dy)
    hr = h.detach().float().requires_grad_(True)
    wr, br = w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    yr = F.layer_norm(F.gelu(hr), (cols,), wr, br, 1e-5)
    yr.backward(dy.float())
    assert rel_l2(y, yr) <= TOL16
    assert rel_l2(h.grad, hr.grad) <= TOL16
    assert rel_l2(w.grad, wr.grad) <= TOL16 and rel_l2(b.grad, br.grad) <= TOL16


@pytest.mark.parametrize("cols", [128, 256, 768, 1024])
def test_ln_res_ln(cols):
    from ofasys_b200 import ops

    gen = g()
    rows = 333
    a = rnd(rows, cols, gen=gen).requires_grad_(True)
    x = rnd(rows, cols, dtype=torch.float32, gen=gen).requires_grad_(True)
    ps = []
    for _ in range(2):
        ps.append((torch.rand(cols, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True))
        ps.append((torch.randn(cols, generator=gen) * 0.1).bfloat16().to(dev()).requires_grad_(True))
    dxn = rnd(rows, cols, dtype=torch.float32, gen=gen)
    dy = rnd(rows, cols, gen=gen)
    xn, y = ops.ln_res_ln(a, x, *ps)
    torch.autograd.backward([xn, y], [dxn, dy])
    ar, xr = a.detach().float().requires_grad_(True), x.detach().clone().requires_grad_(True)
    pr = [p.detach().float().requires_grad_(True) for p in ps]
    xnr = xr + F.layer_norm(ar, (cols,), pr[0], pr[1], 1e-5)
    yr = F.layer_norm(xnr, (cols,), pr[2], pr[3], 1e-5)
    torch.autograd.backward([xnr, yr], [dxn, dy.float()])
    assert rel_l2(xn, xnr) <= TOL32 and rel_l2(y, yr) <= TOL16
    assert rel_l2(x.grad, xr.grad) <= 1e-4 and rel_l2(a.grad, ar.grad) <= TOL16
    for p, q in zip(ps, pr):
        assert rel_l2(p.grad, q.grad) <= TOL16


def test_res_ln_no_first_norm():
    """x_new = x + a ; y = LN(x_new): the deferred FFN residual add fused with the next pre-LN."""
    from ofasys_b200 import ops

    gen = g()
    rows, cols = 301, 768
    a = rnd(rows, cols, gen=gen).requires_grad_(True)
    x = rnd(rows, cols, dtype=torch.float32, gen=gen).requires_grad_(True)
    w = (torch.rand(cols, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    b = (torch.randn(cols, generator=gen) * 0.1).bfloat16().to(dev()).requires_grad_(True)
    dxn, dy = rnd(rows, cols, dtype=torch.float32, gen=gen), rnd(rows, cols, gen=gen)
    xn, y = ops.ln_res_ln(a, x, None, None, w, b)
    torch.autograd.backward([xn, y], [dxn, dy])
    ar, xr, wr, br = a.detach().float().requires_grad_(True), x.detach().clone().requires_grad_(True), w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    xnr = xr + ar
    yr = F.layer_norm(xnr, (cols,), wr, br, 1e-5)
    torch.autograd.backward([xnr, yr], [dxn, dy.float()])
    assert rel_l2(xn, xnr) <= TOL32 and rel_l2(y, yr) <= TOL16
    assert rel_l2(x.grad, xr.grad) <= 1e-4 and rel_l2(a.grad, ar.grad) <= TOL16
    assert rel_l2(w.grad, wr.grad) <= TOL16 and rel_l2(b.grad, br.grad) <= TOL16


def test_colsum():
    from ofasys_b200 import ops

    x = rnd(5000, 200, dtype=torch.float32)
    assert rel_l2(ops.colsum(x, torch.float32), x.sum(0)) <= 1e-5
    xb = rnd(777, 96)
    assert rel_l2(ops.colsum(xb, torch.float32), xb.float().sum(0)) <= 1e-5


# --------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 136, 72), (1000, 768, 768), (257, 2304, 768), (520, 3072, 128), (64, 50265 // 8 * 8 + 8, 256), (768, 768, 1000)])
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    """All four operand-major combinations (fwd: K/K, dgrad: K/MN, wgrad: MN/MN), ragged M/N/K tiles."""
    from ofasys_b200 import ops

    if a_mn and M % 8:
        M = (M + 7) // 8 * 8
    if b_mn and N % 8:
        N = (N + 7) // 8 * 8
    gen = g()
    A = rnd(M, K, gen=gen)
    B = rnd(N, K, gen=gen)
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, dtype=torch.float32, device=dev())
    ops.gemm(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out, N)
    ref = A.float() @ B.float().t()
    assert rel_l2(out, ref) <= TOL32, (M, N, K, a_mn, b_mn)


@pytest.mark.parametrize("bn,cg,cl,bk", [(64, 1, 1, 64), (128, 1, 1, 64), (128, 2, 2, 128), (128, 2, 4, 128), (256, 1, 1, 64), (256, 2, 2, 128),
                                         (256, 2, 4, 128), (128, 2, 2, 64), (256, 2, 2, 64)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_forced_tile_configs(monkeypatch, a_mn, b_mn, bn, cg, cl, bk):
    """Every (BN, cta_group, cluster) instantiation, not only the one the cost model picks: 1-CTA 128xBN tiles,
    CTA-pair 256xBN tiles (tcgen05 cta_group::2, B tile split across the pair) and 4-CTA clusters of two pairs that
    share A by TMA multicast (opt-in through OFAB_GEMM_CL=4, taken when the N-tile count is even)."""
    from ofasys_b200 import ops

    monkeypatch.setenv("OFAB_GEMM_BN", str(bn))
    monkeypatch.setenv("OFAB_GEMM_CG", str(cg))
    monkeypatch.setenv("OFAB_GEMM_CL", str(cl))
    monkeypatch.setenv("OFAB_GEMM_BK", str(bk))  # CTA pairs: 64-deep stages (six, the default) or three 128-deep ones
    gen = g()
    for (M, N, K) in [(1000, 776, 200), (8480 // 4, 2304, 768), (300, 136, 3072), (2048 + 24, 1024, 520)]:
        if a_mn:
            M = (M + 7) // 8 * 8
        if b_mn:
            N = (N + 7) // 8 * 8
        A, B = rnd(M, K, gen=gen), rnd(N, K, gen=gen)
        Am = A.t().contiguous() if a_mn else A
        Bm = B.t().contiguous() if b_mn else B
        bias = rnd(N, gen=gen)
        out = torch.empty(M, N, dtype=torch.float32, device=dev())
        ops.gemm(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out, N, bias=bias)
        ref = A.float() @ B.float().t() + bias.float()
        assert rel_l2(out, ref) <= TOL32, (M, N, K, a_mn, b_mn, bn, cg)
        if N % 8 == 0:  # bf16 output through the swizzled-smem + TMA-store epilogue
            outb = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
            ops.gemm(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, outb, N, bias=bias)
            assert rel_l2(outb.float(), ref) <= 4e-3, (M, N, K, a_mn, b_mn, bn, cg)


def test_gemm_epilogues():
    from ofasys_b200 import ops

    gen = g()
    M, N, K = 300, 776, 192
    A, B = rnd(M, K, gen=gen), rnd(N, K, gen=gen)
    bias = rnd(N, gen=gen)
    res = rnd(M, N, dtype=torch.float32, gen=gen)
    ref = A.float() @ B.float().t()
    o1 = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    ops.gemm(M, N, K, A, K, 0, B, K, 0, o1, N, bias=bias)
    assert rel_l2(o1, ref + bias.float()) <= TOL16
    o2 = torch.empty(M, N, dtype=torch.float32, device=dev())
    ops.gemm(M, N, K, A, K, 0, B, K, 0, o2, N, bias=bias, residual=res, ldr=N)
    assert rel_l2(o2, ref + bias.float() + res) <= TOL32
    # padded leading dimension (logits layout)
    o3 = torch.full((M, N + 8), 7.0, dtype=torch.bfloat16, device=dev())
    ops.gemm(M, N - 3, K, A, K, 0, B, K, 0, o3, N + 8)
    assert rel_l2(o3[:, : N - 3], ref[:, : N - 3]) <= TOL16
    assert (o3[:, N - 3:] == 7.0).all(), "columns beyond N must not be written"


def test_gemm_many_tiles_persistent():
    """more tiles than SMs: exercises the persistent loop, smem ring wrap and both TMEM stages."""
    from ofasys_b200 import ops

    gen = g()
    M, N, K = 4096 + 40, 2304, 768
    A, B = rnd(M, K, gen=gen, scale=0.5), rnd(N, K, gen=gen, scale=0.5)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    ops.gemm(M, N, K, A, K, 0, B, K, 0, out, N)
    assert rel_l2(out, A.float() @ B.float().t()) <= TOL16


@pytest.mark.parametrize("M,N,K,splits", [(768, 768, 8480, 0), (1536, 768, 8480, 0), (3072, 768, 8480, 0), (136, 72, 1000, 0),
                                           (768, 776, 4100, 3), (2304, 768, 2048, 4), (300, 136, 1100, 8), (768, 768, 8480, 6)])
@pytest.mark.parametrize("a_mn,b_mn", [(1, 1), (0, 0)])
def test_gemm_splitk(monkeypatch, M, N, K, splits, a_mn, b_mn):
    """Split-K weight-gradient GEMM (few output tiles, long contraction): partial fp32 slabs + reduction must equal the
    plain product; ragged K tails and K ranges of unequal length included.  splits = 0: the library's own plan
    (768x768 and 1536x768 wgrads of the benchmark step split, 3072x768 and tiny problems do not); > 0: forced."""
    from ofasys_b200 import _lib, ops

    if splits:
        monkeypatch.setenv("OFAB_GEMM_SPLITS", str(splits))
    gen = g()
    if not a_mn:
        K = (K + 7) // 8 * 8  # K-major operands need 16-byte rows
    else:
        M = (M + 7) // 8 * 8  # ... MN-major ones 16-byte columns
    A, B = rnd(M, K, gen=gen, scale=0.5), rnd(N, K, gen=gen, scale=0.5)
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    n_ws = _lib.lib().ofab_gemm_splitk_workspace_elems(M, N, K)
    assert n_ws % (M * N) == 0
    if not splits:
        assert (n_ws > 0) == ((M, N) in ((768, 768), (1536, 768))), (M, N, K, n_ws)
    elif splits > 1:
        nkb = (K + 127) // 128
        per = -(-nkb // min(splits, nkb))  # k-blocks per range; ranges = ceil(nkb / per): no empty range
        assert n_ws == -(-nkb // per) * M * N
    ref = A.float() @ B.float().t()
    out32 = torch.full((M, N), 7.0, dtype=torch.float32, device=dev())
    ops.gemm_splitk(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out32, N)
    assert rel_l2(out32, ref) <= TOL32, (M, N, K)
    out16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev())
    ops.gemm_splitk(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out16, N)
    assert rel_l2(out16, ref) <= 4e-3, (M, N, K)


def test_gemm_first_cuda_call_of_a_fresh_thread():
    """The tensor-map encoder is a driver entry point: a thread that has not bound the primary context yet (a fresh
    autograd worker whose first CUDA work is a GEMM) must still be able to launch."""
    import threading

    from ofasys_b200 import ops

    gen = g()
    M, N, K = 588, 576, 48
    A, B = rnd(M, K, gen=gen), rnd(N, K, gen=gen)
    out = torch.empty(M, N, dtype=torch.float32, device=dev())
    torch.cuda.synchronize()
    err = []

    def work():
        try:
            ops.gemm(M, N, K, A, K, 0, B, K, 0, out, N)
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            err.append(ex)

    th = threading.Thread(target=work)
    th.start()
    th.join()
    assert not err, err
    assert rel_l2(out, A.float() @ B.float().t()) <= TOL32


@pytest.mark.parametrize("with_res", [False, True])
def test_linear_autograd(with_res):
    from ofasys_b200 import ops

    gen = g()
    Bz, T, K, N = 3, 50, 256, 384
    x = rnd(Bz, T, K, gen=gen).requires_grad_(True)
    W = rnd(N, K, gen=gen, scale=0.05).requires_grad_(True)
    b = rnd(N, gen=gen, scale=0.1).requires_grad_(True)
    res = rnd(Bz, T, N, dtype=torch.float32, gen=gen).requires_grad_(True) if with_res else None
    dy = rnd(Bz, T, N, dtype=torch.float32 if with_res else torch.bfloat16, gen=gen)
    y = ops.linear(x, W, b, residual=res)
    y.backward(dy)
    xr, Wr, br = (t.detach().float().requires_grad_(True) for t in (x, W, b))
    yr = F.linear(xr, Wr, br)
    if with_res:
        rr = res.detach().clone().requires_grad_(True)
        yr = yr + rr
    dyr = dy.float() if not with_res else dy.bfloat16().float()  # backward GEMMs consume bf16 dY
    yr.backward(dyr)
    assert rel_l2(y, yr) <= (TOL32 if with_res else TOL16)
    assert rel_l2(x.grad, xr.grad) <= TOL16 and rel_l2(W.grad, Wr.grad) <= TOL16 and rel_l2(b.grad, br.grad) <= TOL16
    if with_res:
        assert torch.equal(res.grad, dy)


# ---------------------------------------------------------------------------------- attention
def attn_ref(q, k, v, pq, pk, table, idx, kpm, causal, scale, H):
    """fp32 restatement of multihead_attention.py:308-338 with the bias written out densely."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    dh = 64
    sp = lambda t: t.view(t.shape[0], t.shape[1], H, dh).transpose(1, 2)
    s = torch.matmul(sp(q), sp(k).transpose(2, 3))
    if pq is not None:
        s = s + torch.matmul(sp(pq), sp(pk).transpose(2, 3))
    s = s * scale
    if idx is not None:
        bias = table[idx.clamp_min(0).long()]  # Tq x Tk x H
        bias = torch.where((idx >= 0).unsqueeze(-1), bias, torch.zeros_like(bias))
        s = s + bias.permute(2, 0, 1).unsqueeze(0)
    if causal:
        s = s + torch.triu(torch.full((Tq, Tk), float("-inf"), device=s.device), 1)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, sp(v)).transpose(1, 2).reshape(B, Tq, d)


@pytest.mark.parametrize("mode", ["self", "cross"])
@pytest.mark.parametrize("Tq,Tk", [(64, 64), (24, 24), (130, 130), (16, 265), (257, 257), (700, 700), (1, 77)])
@pytest.mark.parametrize("variant", ["plain", "pos", "pos_rel", "pos_rel_kpm_causal", "kpm", "causal"])
def test_attention(mode, Tq, Tk, variant):
    from ofasys_b200 import ops

    if mode == "self" and Tq != Tk:
        pytest.skip("self-attention needs Tq == Tk")
    if mode == "cross" and "causal" in variant:
        pytest.skip("no causal cross-attention")
    gen = g()
    B, H = 2, 2
    d = H * 64
    scale = 128 ** -0.5
    if mode == "self":
        qkv = rnd(B, Tq, 3 * d, gen=gen).requires_grad_(True)
        kv = None
    else:
        qkv = rnd(B, Tq, d, gen=gen).requires_grad_(True)
        kv = rnd(B, Tk, 2 * d, gen=gen).requires_grad_(True)
    pq = pk = table = idx = kpm = None
    if "pos" in variant:
        pq = rnd(1, Tq, d, gen=gen).requires_grad_(True)
        pk = rnd(1, Tk, d, gen=gen).requires_grad_(True)
    if "rel" in variant:
        nb = 37
        table = rnd(nb, H, gen=gen, scale=0.5).requires_grad_(True)
        idx = torch.randint(-1, nb, (Tq, Tk), generator=gen).to(torch.int32).to(dev())
    if "kpm" in variant:
        kpm = torch.zeros(B, Tk, dtype=torch.bool)
        kpm[1, Tk - max(1, Tk // 4):] = True
        if Tk > 8:
            kpm[0, 3] = True  # interior pad (concatenated slots)
        kpm = kpm.to(dev())
    causal = "causal" in variant
    do = rnd(B, Tq, d, gen=gen)
    o = ops.attention(qkv, kv, H, scale, ops.PositionBias(pq, pk, idx, table), kpm, causal)
    o.backward(do)

    leaves = [t for t in (qkv, kv, pq, pk, table) if t is not None]
    refs = {id(t): t.detach().float().requires_grad_(True) for t in leaves}
    R = lambda t: None if t is None else refs[id(t)]
    if mode == "self":
        q_, k_, v_ = R(qkv)[..., :d], R(qkv)[..., d:2 * d], R(qkv)[..., 2 * d:]
    else:
        q_, k_, v_ = R(qkv), R(kv)[..., :d], R(kv)[..., d:]
    pq_ = None if pq is None else R(pq).expand(B, -1, -1)
    pk_ = None if pk is None else R(pk).expand(B, -1, -1)
    orf = attn_ref(q_, k_, v_, pq_, pk_, R(table), idx, kpm, causal, scale, H)
    orf.backward(do.float())
    assert rel_l2(o, orf) <= 1e-2, (mode, Tq, Tk, variant)
    for t in leaves:
        assert rel_l2(t.grad, refs[id(t)].grad) <= 2e-2, (mode, Tq, Tk, variant, tuple(t.shape))


@pytest.mark.parametrize("mode", ["self", "cross"])
@pytest.mark.parametrize("Tq,Tk", [(64, 64), (24, 24), (130, 130), (16, 265), (257, 257), (520, 520), (200, 700)])
@pytest.mark.parametrize("variant", ["pos", "pos_rel", "pos_rel_kpm_causal", "pos_kpm"])
def test_attention_dense_bias(mode, Tq, Tk, variant):
    """Mode A on the tcgen05 kernels: the position terms enter as ONE dense batch-invariant tile (abs_pos + table gather);
    the gradients of pq / pk / table come back through the batch-reduced dS."""
    from ofasys_b200 import ops

    if mode == "self" and Tq != Tk:
        pytest.skip("self-attention needs Tq == Tk")
    if mode == "cross" and "causal" in variant:
        pytest.skip("no causal cross-attention")
    gen = g()
    B, H = 3, 2
    d = H * 64
    scale = 128 ** -0.5
    if mode == "self":
        qkv = rnd(B, Tq, 3 * d, gen=gen).requires_grad_(True)
        kv = None
    else:
        qkv = rnd(B, Tq, d, gen=gen).requires_grad_(True)
        kv = rnd(B, Tk, 2 * d, gen=gen).requires_grad_(True)
    table = idx = kpm = None
    pq = rnd(1, Tq, d, gen=gen).requires_grad_(True)
    pk = rnd(1, Tk, d, gen=gen).requires_grad_(True)
    if "rel" in variant:
        nb = 37
        table = rnd(nb, H, gen=gen, scale=0.5).requires_grad_(True)
        idx = torch.randint(-1, nb, (Tq, Tk), generator=gen).to(torch.int32).to(dev())
    if "kpm" in variant:
        kpm = torch.zeros(B, Tk, dtype=torch.bool)
        kpm[1, Tk - max(1, Tk // 4):] = True
        if Tk > 8:
            kpm[0, 3] = True
        kpm = kpm.to(dev())
    causal = "causal" in variant
    do = rnd(B, Tq, d, gen=gen)
    o = ops.attention(qkv, kv, H, scale, ops.PositionBias(pq, pk, idx, table, abs=ops.abs_pos(pq, pk, H)), kpm, causal)
    o.backward(do)
    leaves = [t for t in (qkv, kv, pq, pk, table) if t is not None]
    refs = {id(t): t.detach().float().requires_grad_(True) for t in leaves}
    R = lambda t: None if t is None else refs[id(t)]
    if mode == "self":
        q_, k_, v_ = R(qkv)[..., :d], R(qkv)[..., d:2 * d], R(qkv)[..., 2 * d:]
    else:
        q_, k_, v_ = R(qkv), R(kv)[..., :d], R(kv)[..., d:]
    orf = attn_ref(q_, k_, v_, R(pq).expand(B, -1, -1), R(pk).expand(B, -1, -1), R(table), idx, kpm, causal, scale, H)
    orf.backward(do.float())
    assert rel_l2(o, orf) <= 1e-2, (mode, Tq, Tk, variant)
    for t in leaves:
        assert rel_l2(t.grad, refs[id(t)].grad) <= 2e-2, (mode, Tq, Tk, variant, tuple(t.shape))


def test_attention_legacy_kernels_agree():
    """OFAB_ATTN_LEGACY keeps the mma.sync kernels reachable for A/B runs; structured position terms (no dense abs) still
    take them (incremental decoding slices): both paths give the same numbers."""
    from ofasys_b200 import ops

    gen = g()
    B, H, T = 2, 2, 130
    d = H * 64
    qkv = rnd(B, T, 3 * d, gen=gen).requires_grad_(True)
    pq = rnd(1, T, d, gen=gen)
    pk = rnd(1, T, d, gen=gen)
    table = rnd(37, H, gen=gen, scale=0.5)
    idx = torch.randint(-1, 37, (T, T), generator=gen).to(torch.int32).to(dev())
    o1 = ops.attention(qkv, None, H, 0.1, ops.PositionBias(pq, pk, idx, table), None, True)  # structured: mma.sync kernels
    o2 = ops.attention(qkv, None, H, 0.1, ops.PositionBias(pq, pk, idx, table, abs=ops.abs_pos(pq, pk, H)), None, True)  # dense: tcgen05
    assert rel_l2(o2, o1) <= 6e-3


# ------------------------------------------------------------------------------- embed / CE
@pytest.mark.parametrize("d", [128, 256, 768])
@pytest.mark.parametrize("entangle,is_src", [(False, True), (True, True), (True, False)])
def test_embed_ln_gather(d, entangle, is_src):
    from ofasys_b200 import ops

    gen = g()
    B, T, V = 3, 19, 97
    E = rnd(V, d, gen=gen, scale=0.5).requires_grad_(True)
    pos = rnd(T + 5, d, gen=gen, scale=0.5).requires_grad_(True)
    typ = rnd(1, d, gen=gen, scale=0.5).requires_grad_(True)
    gam = (torch.rand(d, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    bet = rnd(d, gen=gen, scale=0.1).requires_grad_(True)
    tok = torch.randint(2, V, (B, T), generator=gen)
    tok[1, -4:] = 1
    tok = tok.to(dev())
    mask = tok.eq(1)
    dout = rnd(B, T, d, dtype=torch.float32, gen=gen)
    out = ops.embed_ln(gam, bet, tokens=tok, E=E, pos=pos if entangle else None, type_vec=typ if is_src else None,
                       zero_mask=mask if is_src else None, padding_idx=1)
    out.backward(dout)
    Er, pr, tr, gr, br = (t.detach().float().requires_grad_(True) for t in (E, pos, typ, gam, bet))
    pre = F.embedding(tok, Er, padding_idx=1)
    if entangle:
        pre = pre + pr[:T]
    if is_src:
        pre = pre + tr.squeeze()
    ref = F.layer_norm(pre, (d,), gr, br, 1e-5)
    if is_src:
        ref = ref * (1 - mask.unsqueeze(-1).float())
    ref.backward(dout)
    assert rel_l2(out, ref) <= TOL32
    assert rel_l2(E.grad, Er.grad) <= TOL16
    assert rel_l2(gam.grad, gr.grad) <= TOL16 and rel_l2(bet.grad, br.grad) <= TOL16
    if entangle:
        assert rel_l2(pos.grad, pr.grad) <= TOL16
    if is_src:
        assert rel_l2(typ.grad, tr.grad) <= TOL16


@pytest.mark.parametrize("has_cls", [False, True])
def test_embed_ln_dense(has_cls):
    from ofasys_b200 import ops

    gen = g()
    B, P, d = 2, 9, 256
    dense = rnd(B, P, d, gen=gen).requires_grad_(True)
    cls = rnd(1, 1, d, gen=gen).requires_grad_(True) if has_cls else None
    T = P + int(has_cls)
    pos = rnd(T, d, gen=gen).requires_grad_(True)
    typ = rnd(1, d, gen=gen).requires_grad_(True)
    gam = (torch.rand(d, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    bet = rnd(d, gen=gen, scale=0.1).requires_grad_(True)
    dout = rnd(B, T, d, dtype=torch.float32, gen=gen)
    out = ops.embed_ln(gam, bet, dense=dense, cls=cls, pos=pos, type_vec=typ)
    out.backward(dout)
    leaves = [t for t in (dense, cls, pos, typ, gam, bet) if t is not None]
    R = {id(t): t.detach().float().requires_grad_(True) for t in leaves}
    x = R[id(dense)]
    if has_cls:
        x = torch.cat([R[id(cls)].expand(B, -1, -1), x], dim=1)
    ref = F.layer_norm(x + R[id(pos)][:T] + R[id(typ)].squeeze(), (d,), R[id(gam)], R[id(bet)], 1e-5)
    ref.backward(dout)
    assert rel_l2(out, ref) <= TOL32
    for t in leaves:
        assert rel_l2(t.grad, R[id(t)].grad) <= TOL16, tuple(t.shape)


@pytest.mark.parametrize("V", [512, 50265])
def test_cross_entropy_and_fused_projection(V):
    from ofasys_b200 import ops

    gen = g()
    M, d = 77, 128
    x = rnd(M, d, gen=gen).requires_grad_(True)
    E = rnd(V, d, gen=gen, scale=0.3).requires_grad_(True)
    tgt = torch.randint(2, V, (M,), generator=gen)
    tgt[::7] = 1
    tgt = tgt.to(dev())
    loss = ops.linear_cross_entropy(x, E, tgt, 1)
    (loss * 0.5).backward()
    xr, Er = x.detach().float().requires_grad_(True), E.detach().float().requires_grad_(True)
    logits = (xr @ Er.t()).bfloat16().float()  # the kernel reads bf16 logits
    lr = F.cross_entropy(logits, tgt, ignore_index=1, reduction="sum")
    lref = F.cross_entropy(xr @ Er.t(), tgt, ignore_index=1, reduction="sum")
    (lref * 0.5).backward()
    assert abs(loss.item() - lr.item()) <= 1e-4 * abs(lr.item())
    assert abs(loss.item() - lref.item()) <= 2e-3 * abs(lref.item())
    assert rel_l2(x.grad, xr.grad) <= 1e-2 and rel_l2(E.grad, Er.grad) <= 1e-2
    # stand-alone criterion on model logits
    lg = ops.linear(x.detach(), E.detach()).requires_grad_(True)
    l2 = ops.cross_entropy_sum(lg, tgt, 1)
    l2.backward()
    lgr = lg.detach().float().requires_grad_(True)
    F.cross_entropy(lgr, tgt, ignore_index=1, reduction="sum").backward()
    assert abs(l2.item() - lr.item()) <= 1e-4 * abs(lr.item())
    assert rel_l2(lg.grad, lgr.grad) <= TOL16


@pytest.mark.parametrize("V,chunk", [(512, 32), (50265, 40)])
def test_chunked_fused_projection_criterion_matches_the_one_shot_form(V, chunk):
    """Row-chunked projection + criterion (no [rows, V] scratch): same loss, nll and gradients as the
    one-shot fused form and as torch fp32, label smoothing on, ragged last chunk."""
    from ofasys_b200 import ops

    gen = g()
    M, d, eps = 77, 128, 0.1
    x0 = rnd(M, d, gen=gen)
    E0 = rnd(V, d, gen=gen, scale=0.3)
    tgt = torch.randint(2, V, (M,), generator=gen)
    tgt[::7] = 1
    tgt = tgt.to(dev())
    out = []
    for ch in (None, chunk):
        x, E = x0.clone().requires_grad_(True), E0.clone().requires_grad_(True)
        nll = torch.zeros(1, device=dev())
        loss = ops.linear_cross_entropy(x, E, tgt, 1, eps, nll, chunk_rows=ch)
        (loss * 0.37).backward()
        out.append((loss.item(), nll.item(), x.grad, E.grad))
    (l1, n1, dx1, dE1), (l2, n2, dx2, dE2) = out
    assert abs(l1 - l2) <= 1e-6 * abs(l1) and abs(n1 - n2) <= 1e-6 * abs(n1)
    assert rel_l2(dx2, dx1) <= 3e-3 and rel_l2(dE2, dE1) <= 3e-3  # same bf16 gradient tile; split-K / fp32-accumulated sums
    from oracle import oracle_model as om

    xr, Er = x0.float().cpu().requires_grad_(True), E0.float().cpu().requires_grad_(True)
    lref, nref, _ = om.label_smoothed_cross_entropy_sum(xr @ Er.t(), tgt.cpu(), eps)  # label_smoothed_cross_entropy.py:62-92
    (lref * 0.37).backward()
    assert abs(l2 - lref.item()) <= 3e-3 * abs(lref.item()) and abs(n2 - nref.item()) <= 3e-3 * abs(nref.item())
    assert rel_l2(dx2, xr.grad) <= 1e-2 and rel_l2(dE2, Er.grad) <= 1e-2


def test_scale_cols():
    from ofasys_b200 import ops

    gen = g()
    W = rnd(128, 128, gen=gen).requires_grad_(True)
    c = (torch.rand(2, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    dW = rnd(128, 128, gen=gen)
    We = ops.scale_cols(W, c, 64)
    We.backward(dW)
    Wr, cr = W.detach().float().requires_grad_(True), c.detach().float().requires_grad_(True)
    (Wr * cr.repeat_interleave(64)[None, :]).backward(dW.float())
    assert rel_l2(We, Wr * cr.repeat_interleave(64)[None, :]) <= TOL16
    assert rel_l2(W.grad, Wr.grad) <= TOL16 and rel_l2(c.grad, cr.grad) <= TOL16


# ------------------------------------------------------------------------------- audio convs
def test_audio_subsampler_pieces():
    from ofasys_b200 import ops
    from ofasys_b200.adaptor.audio import Conv2dSubsampling4

    gen = g()
    B, L, Fd, C = 2, 61, 80, 64
    sub = Conv2dSubsampling4(Fd, C).to(dev())
    with torch.no_grad():
        for p in sub.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * (0.3 if p.dim() > 1 else 0.1))
    ref = [p.detach().clone().float().requires_grad_(True) for p in sub.parameters()]
    sub = sub.to(torch.bfloat16)
    for p, r in zip(sub.parameters(), ref):
        r.data.copy_(p.detach().float())
    x = rnd(B, L, Fd, dtype=torch.float32, gen=gen)
    lens = torch.tensor([L, L - 9], device=dev())
    out, ol = sub(x, lens)
    dy = rnd(*out.shape, gen=gen)
    out.backward(dy)
    w1, b1, w2, b2, wl, bl = ref
    h = F.relu(F.conv2d(x.unsqueeze(1), w1, b1, stride=2)).bfloat16().float()
    h = F.relu(F.conv2d(h, w2, b2, stride=2)).bfloat16().float()
    b_, c_, t_, f_ = h.shape
    yr = F.linear(h.transpose(1, 2).contiguous().view(b_, t_, c_ * f_), wl, bl)
    yr.backward(dy.float())
    assert tuple(out.shape) == tuple(yr.shape)
    assert rel_l2(out, yr) <= 1e-2
    for p, r in zip(sub.parameters(), ref):
        assert rel_l2(p.grad, r.grad) <= 3e-2, tuple(p.shape)


# ------------------------------------------------------------------------------- ResNet pieces
@pytest.mark.parametrize("k,stride,pad,H,W,C", [(3, 1, 1, 14, 14, 64), (3, 2, 1, 15, 13, 32), (3, 2, 1, 56, 56, 64)])
def test_conv_kxk_im2col_gemm(k, stride, pad, H, W, C):
    """k x k convolution = im2col (channel-last) + tcgen05 GEMM with the [Co,Ci,k*k] -> [Co,k*k,Ci] weight permute."""
    from ofasys_b200 import ops

    gen = g()
    B, Co = 3, 48
    x = rnd(B, H, W, C, gen=gen).requires_grad_(True)
    w = rnd(Co, C, k, k, gen=gen, scale=0.05).requires_grad_(True)
    cols = ops.im2col_nhwc(x, k, stride, pad)
    wp = ops.transpose_last2(w.reshape(Co, C, k * k)).reshape(Co, k * k * C)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = ops.linear(cols, wp, None).view(B, Ho, Wo, Co)
    dy = rnd(B, Ho, Wo, Co, gen=gen)
    y.backward(dy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=pad)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(y.permute(0, 3, 1, 2), yr) <= TOL16
    assert rel_l2(x.grad.permute(0, 3, 1, 2), xr.grad) <= 1e-2
    assert rel_l2(w.grad, wr.grad) <= 1e-2


def test_stem_maxpool_subsample():
    from ofasys_b200 import ops

    gen = g()
    B = 2
    img = rnd(B, 3, 64, 48, dtype=torch.float32, gen=gen)
    w = rnd(64, 3, 7, 7, gen=gen, scale=0.1).requires_grad_(True)
    cols, Ho, Wo = ops.im2col_nchw(img, 7, 2, 3, 152)
    y = ops.linear(cols, F.pad(w.reshape(64, 147), (0, 5)), None).view(B, Ho, Wo, 64)
    yr = F.conv2d(img.bfloat16().float(), w.detach().float(), stride=2, padding=3)
    assert rel_l2(y.permute(0, 3, 1, 2), yr) <= TOL16
    # maxpool
    x = rnd(B, 17, 15, 64, gen=gen).requires_grad_(True)
    p = ops.maxpool3x3s2(x)
    dp = rnd(*p.shape, gen=gen)
    p.backward(dp)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    pr = F.max_pool2d(xr, 3, 2, 1)
    pr.backward(dp.float().permute(0, 3, 1, 2))
    assert torch.equal(p.permute(0, 3, 1, 2).float(), pr)
    assert rel_l2(x.grad.permute(0, 3, 1, 2), xr.grad) <= TOL16
    # stride-2 subsample
    x2 = rnd(B, 9, 8, 32, gen=gen).requires_grad_(True)
    s = ops.subsample2(x2)
    assert torch.equal(s, x2[:, ::2, ::2])
    ds = rnd(*s.shape, gen=gen)
    s.backward(ds)
    ref = torch.zeros_like(x2)
    ref[:, ::2, ::2] = ds
    assert torch.equal(x2.grad, ref)


@pytest.mark.parametrize("relu,with_res", [(False, False), (True, False), (True, True)])
def test_batch_norm_train(relu, with_res):
    from ofasys_b200 import ops

    gen = g()
    B, H, W, C = 4, 7, 9, 64
    x = (rnd(B, H, W, C, gen=gen).float() * 2 + 0.7).bfloat16().requires_grad_(True)
    res = rnd(B, H, W, C, gen=gen).requires_grad_(True) if with_res else None
    gam = (torch.rand(C, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(True)
    bet = rnd(C, gen=gen, scale=0.1).requires_grad_(True)
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    y = ops.batch_norm_train(x, gam, bet, res, relu, 1e-5, 0.1, rm, rv)
    dy = rnd(B, H, W, C, gen=gen)
    y.backward(dy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gam.detach().float().requires_grad_(True), bet.detach().float().requires_grad_(True)
    rmr, rvr = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    yr = F.batch_norm(xr, rmr, rvr, gr, br, True, 0.1, 1e-5)
    if with_res:
        rr = res.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(y.permute(0, 3, 1, 2), yr) <= TOL16
    assert rel_l2(x.grad.permute(0, 3, 1, 2), xr.grad) <= 2e-2
    assert rel_l2(gam.grad, gr.grad) <= 2e-2 and rel_l2(bet.grad, br.grad) <= 2e-2
    assert rel_l2(rm, rmr) <= 1e-4 and rel_l2(rv, rvr) <= 1e-3
    if with_res:
        assert rel_l2(res.grad.permute(0, 3, 1, 2), rr.grad) <= TOL16


@pytest.mark.parametrize("relu,with_res,frozen", [(True, False, False), (True, True, False), (False, False, True), (True, True, True)])
def test_batch_norm_eval_and_frozen(relu, with_res, frozen):
    """nn.BatchNorm2d in eval mode (model.eval(), or `freeze_resnet`: adaptor/image_resnet.py:107-114 -- BatchNorm in eval mode
    with frozen affine parameters while the convolutions keep training) against F.batch_norm(training=False): the running
    statistics are used and NOT updated, dx = gamma * rstd * g, the affine gradients only when they are trainable."""
    from ofasys_b200 import ops

    gen = g()
    B, H, W, C = 4, 7, 9, 64
    x = (rnd(B, H, W, C, gen=gen).float() * 2 + 0.7).bfloat16().requires_grad_(True)
    res = rnd(B, H, W, C, gen=gen).requires_grad_(True) if with_res else None
    gam = (torch.rand(C, generator=gen) + 0.5).bfloat16().to(dev()).requires_grad_(not frozen)
    bet = rnd(C, gen=gen, scale=0.1).requires_grad_(not frozen)
    rm = (torch.randn(C, generator=gen) * 0.3 + 0.5).to(dev())
    rv = (torch.rand(C, generator=gen) * 2 + 0.5).to(dev())
    rm0, rv0 = rm.clone(), rv.clone()
    y = ops.batch_norm_eval(x, gam, bet, rm, rv, res, relu, 1e-5)
    dy = rnd(B, H, W, C, gen=gen)
    y.backward(dy)
    assert torch.equal(rm, rm0) and torch.equal(rv, rv0)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    gr, br = gam.detach().float().requires_grad_(True), bet.detach().float().requires_grad_(True)
    yr = F.batch_norm(xr, rm0.clone(), rv0.clone(), gr, br, False, 0.1, 1e-5)
    if with_res:
        rr = res.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_l2(y.permute(0, 3, 1, 2), yr) <= TOL16
    assert rel_l2(x.grad.permute(0, 3, 1, 2), xr.grad) <= 2e-2
    if frozen:
        assert gam.grad is None and bet.grad is None
    else:
        assert rel_l2(gam.grad, gr.grad) <= 2e-2 and rel_l2(bet.grad, br.grad) <= 2e-2
    if with_res:
        assert rel_l2(res.grad.permute(0, 3, 1, 2), rr.grad) <= TOL16


def test_freeze_resnet_adaptor_matches_the_oracle_in_eval_statistics():
    """`freeze_resnet=True` (docs/source/howto/train.rst:66-69 fine-tuning recipe): after model.train() the backbone's BatchNorms
    are in eval mode with frozen affine parameters; the forward equals the oracle's backbone run with training=False and the
    convolution weights still receive gradients."""
    from ofasys_b200.module.resnet import resnet50_backbone
    from oracle import oracle_model as om
    import torch.nn as nn

    gen = g()
    net = resnet50_backbone().to(dev())
    with torch.no_grad():
        for n, p in net.named_parameters():
            if p.dim() == 4:
                p.copy_(torch.randn(p.shape, generator=gen) * (2.0 / (p.shape[1] * p.shape[2] * p.shape[3])) ** 0.5)
        for m_ in net.modules():
            if isinstance(m_, nn.BatchNorm2d):
                m_.running_mean.copy_(torch.randn(m_.running_mean.shape, generator=gen) * 0.1)
                m_.running_var.copy_(torch.rand(m_.running_var.shape, generator=gen) + 0.5)
                m_.weight.copy_(torch.rand(m_.weight.shape, generator=gen) * 0.5 + 0.5)
    sd = {"r." + k: v.detach().float().cpu() for k, v in net.state_dict().items() if v.is_floating_point()}
    net = net.bfloat16()
    net.train()
    for m_ in net.modules():  # what ImageResnetAdaptor.train() does with freeze_resnet
        if isinstance(m_, nn.BatchNorm2d):
            m_.eval()
            m_.weight.requires_grad = False
            m_.bias.requires_grad = False
    x = torch.randn(2, 3, 64, 64, generator=gen).to(dev())
    y = net(x)  # [B, h, w, 1024]
    y.float().sum().backward()
    sd16 = {k: v.bfloat16().float() for k, v in sd.items()}
    yr = om.resnet_backbone(x.cpu().bfloat16().float(), sd16, "r", "resnet50", training=False)
    assert rel_l2(y.permute(0, 3, 1, 2), yr) <= 3e-2
    assert net.conv1.weight.grad is not None and net.layer3[0].conv2.weight.grad.abs().sum() > 0
    assert all(m_.weight.grad is None for m_ in net.modules() if isinstance(m_, nn.BatchNorm2d))
    nbt = [int(m_.num_batches_tracked) for m_ in net.modules() if isinstance(m_, nn.BatchNorm2d)]
    assert all(n == 0 for n in nbt)  # eval-mode BatchNorms do not count batches


def test_resnet50_backbone_vs_oracle():
    """whole C1..C4 backbone (train-mode BN) against the oracle's restatement run on the GPU in fp32."""
    from ofasys_b200.module.resnet import resnet50_backbone
    from oracle import oracle_model as om

    gen = g()
    net = resnet50_backbone().to(dev())
    with torch.no_grad():
        for n, p in net.named_parameters():
            if p.dim() == 4:
                fan = p.shape[1] * p.shape[2] * p.shape[3]
                p.copy_(torch.randn(p.shape, generator=gen) * (2.0 / fan) ** 0.5)
            elif n.endswith("bn3.weight"):  # small residual-branch gain keeps the random net out of the chaotic regime
                p.copy_(torch.rand(p.shape, generator=gen) * 0.2 + 0.1)
            elif n.endswith("weight"):
                p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    net = net.to(torch.bfloat16).train()
    sd = {"r." + k: v.detach().float() for k, v in net.state_dict().items() if v.is_floating_point()}
    ref_leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "running" not in k}
    sd_ref = dict(sd)
    sd_ref.update(ref_leaves)
    img = rnd(4, 3, 64, 64, dtype=torch.float32, gen=gen)
    y = net(img)
    dy = rnd(*y.shape, gen=gen)
    y.backward(dy)
    # fp32 truth, and the same algorithm with bf16 activation storage (fp32 arithmetic) as the yardstick: ReLU-mask
    # flips make bf16-storage gradients differ by tens of percent from an fp32 forward (see tests/test_model_gpu.py)
    yr = om.resnet_backbone(img, sd_ref, "r", "resnet50", training=True)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    sim_leaves = {k: v.detach().clone().requires_grad_(True) for k, v in ref_leaves.items()}
    sd_sim = dict(sd)
    sd_sim.update(sim_leaves)
    om.STORE_BF16 = True
    try:
        ys = om.resnet_backbone(img, sd_sim, "r", "resnet50", training=True)
        ys.backward(dy.float().permute(0, 3, 1, 2))
    finally:
        om.STORE_BF16 = False
    e, es = rel_l2(y.permute(0, 3, 1, 2), yr), rel_l2(ys, yr)
    assert e <= max(1.5 * es, 1e-2), (e, es)
    n1 = n2 = den = 0.0
    for k, p in net.named_parameters():
        gr = ref_leaves["r." + k].grad
        n1 += (p.grad.float() - gr).pow(2).sum().item()
        n2 += (sim_leaves["r." + k].grad - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    ours, ref16 = (n1 / den) ** 0.5, (n2 / den) ** 0.5
    assert ours <= 1.5 * ref16 + 1e-2, (ours, ref16)


def test_multi_copy_pack_unpack():
    """flat-bucket pack / unpack of odd-sized gradient tensors (ofab_multi_copy) is a bit-exact round trip."""
    from ofasys_b200 import _lib
    from ofasys_b200.distributed import GradBuckets

    gen = g()
    shapes = [(768,), (3, 5, 7), (1000, 768), (1,), (257, 33), (131072 + 3,)]
    params = [torch.nn.Parameter(rnd(*s, gen=gen)) for s in shapes]
    for p in params:
        p.grad = rnd(*p.shape, gen=gen)
    ref = [p.grad.clone() for p in params]
    gb = GradBuckets(params, bucket_bytes=1 << 20)
    st = torch.cuda.current_stream().cuda_stream
    for i, bucket in enumerate(gb.buckets):
        grads = [p.grad for p in bucket]
        tp, tu = gb._tables(i, grads)
        _lib.call("ofab_multi_copy", tp.data_ptr(), tp.shape[0], st)
        flat = gb._flat[i]
        off = 0
        for gr in grads:  # packed at 16-byte aligned offsets, bit-exact
            assert torch.equal(flat[off:off + gr.numel()].view_as(gr), gr)
            off += (gr.numel() * 2 + 15) // 16 * 8
        flat.mul_(2.0)
        _lib.call("ofab_multi_copy", tu.data_ptr(), tu.shape[0], st)
    for p, r in zip(params, ref):
        assert torch.equal(p.grad, (r.float() * 2).bfloat16())


def test_dp_nccl_two_gpus():
    """world-size-2 NCCL run of the data-parallel exchange (skipped on a one-GPU box)."""
    import subprocess, sys, os

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tests", "dp_nccl_worker.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_NCCL_OK" in r.stdout


def test_dp_model_equality():
    """Model-level data-parallel gate: N-rank gradients after the exchange and the trainer's rescale == one
    process on the concatenated batch, incl. an adaptor no batch touches.  Uses every visible GPU (a 1-GPU box runs world
    size 1: the exchange is then the identity and the check covers the wrapper + rescale arithmetic)."""
    import subprocess, sys, os

    n = max(1, min(torch.cuda.device_count(), 8))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29537", os.path.join(root, "tests", "dp_model_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_MODEL_OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_video_frames_and_zero_mask(dtype):
    """clip -> frames transpose + the reference's all-zero-frame padding test (bit-exact)."""
    from ofasys_b200 import ops

    gen = g()
    B, C, F_, H, W = 3, 3, 5, 17, 13
    clip = torch.randn(B, C, F_, H, W, generator=gen).to(dev(), dtype)
    clip[1, :, 2] = 0
    clip[2, :, 4] = 0
    clip[2, 1, 4, 16, 12] = -0.0
    clip[0, :, 0] = 0
    clip[0, 2, 0, 3, 3] = 1e-30 if dtype == torch.float32 else 1e-20  # one tiny non-zero value: not padding
    frames, zero = ops.video_frames(clip)
    v = clip.transpose(1, 2)
    ref_mask = v.reshape(B, F_, -1).float().abs().mean(dim=-1) == 0.0
    tiny = clip[0, :, 0].float().abs().sum() > 0
    assert tiny
    want = ref_mask.clone()
    want[0, 0] = False  # fp32 mean of one 1e-30 among 663 zeros underflows in the reference formula only below 1e-38
    assert torch.equal(zero, want)
    assert torch.equal(frames, v.reshape(B * F_, C, H, W).to(torch.bfloat16))


# ------------------------------------------------------------------- programmatic dependent launch
def test_pdl_chain_matches_plain_stream_order():
    """Kernels launched with programmatic stream serialization start before their predecessor has finished and wait
    in griddepcontrol.wait before touching global memory: a dependent chain (GEMM -> LN junction -> GELU+LN -> GEMM,
    fwd + bwd, eager and replayed from a CUDA graph) must give bit-identical results with the attribute on and off."""
    from ofasys_b200 import _lib, ops

    gen = g()
    rows, d, f = 1061, 256, 1024
    x0 = rnd(rows, d, dtype=torch.float32, gen=gen)
    w1, b1 = rnd(f, d, gen=gen, scale=0.05), rnd(f, gen=gen, scale=0.1)
    w2, b2 = rnd(d, f, gen=gen, scale=0.05), rnd(d, gen=gen, scale=0.1)
    lw = [(torch.rand(n, generator=gen) + 0.5).bfloat16().to(dev()) for n in (d, d, f)]
    lb = [rnd(n, gen=gen, scale=0.1) for n in (d, d, f)]
    leaves = [w1, b1, w2, b2] + lw + lb
    for t in leaves:
        t.requires_grad_(True)

    def chain():
        for t in leaves:
            t.grad = None
        x = x0
        a = ops.cast_bf16(x0)
        for _ in range(6):
            x, y = ops.ln_res_ln(a, x, lw[0], lb[0], lw[1], lb[1], 1e-5)
            h = ops.linear(y, w1, b1)
            h = ops.layer_norm(h, lw[2], lb[2], 1e-5, gelu=True)
            a = ops.linear(h, w2, b2)
        out = ops.add_residual(x, a)
        out.square().sum().backward()
        return [out.detach().clone()] + [t.grad.detach().clone() for t in leaves]

    lib = _lib.lib()
    prev = lib.ofab_set_pdl(0)
    try:
        ref = chain()
        torch.cuda.synchronize()
        lib.ofab_set_pdl(1)
        for _ in range(3):
            got = chain()
            torch.cuda.synchronize()
            for r, t in zip(ref, got):
                assert torch.equal(r, t)
        # the same chain captured with programmatic edges and replayed
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            chain()
        torch.cuda.current_stream().wait_stream(s)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            outs = chain()
        for _ in range(3):
            gr.replay()
            torch.cuda.synchronize()
            for r, t in zip(ref, outs):
                assert torch.equal(r, t)
    finally:
        lib.ofab_set_pdl(prev)


# ------------------------------------------------------------------------------ label-smoothed criterion
def test_label_smoothed_cross_entropy_vs_oracle_and_reference_fixture():
    """ofab_ce_fwd / _bwd with label_smoothing > 0 against the oracle restatement and the reference's
    own label_smoothed_nll_loss outputs (tests/golden/ls_ce.pt); V = 1003 exercises the ragged tail (ld padded to 1008).
    Loss in fp32 -> 2e-5 relative; dlogits are stored in bf16 -> 6e-3 rel-L2."""
    import os

    from ofasys_b200 import ops
    from oracle import oracle_model as om

    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ls_ce.pt"), weights_only=False)
    logits, target, eps = om.make_ls_case()
    x = logits.to(dev()).requires_grad_(True)
    loss = ops.cross_entropy_sum(x, target.to(dev()), ignore_index=1, label_smoothing=eps)
    loss.backward()
    xr = logits.float().requires_grad_(True)
    lr, nr, _ = om.label_smoothed_cross_entropy_sum(xr, target, eps)
    lr.backward()
    assert abs(loss.item() - lr.item()) <= 2e-5 * abs(lr.item())
    assert abs(loss.item() - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"]))
    assert rel_l2(x.grad, xr.grad) <= 6e-3
    assert rel_l2(x.grad, fx["dlogits"]) <= 6e-3
    pad_rows = (target == 1).nonzero().flatten()
    assert not x.grad[pad_rows.to(dev())].any()  # padding rows: exactly zero gradient


def test_linear_label_smoothed_cross_entropy_fused():
    """The tied projection fused with the label-smoothed criterion, plus the plain nll for logging."""
    from ofasys_b200 import ops
    from oracle import oracle_model as om

    gen = g()
    M, K, V, eps = 48, 128, 515, 0.1
    x = rnd(M, K, gen=gen, scale=0.5).requires_grad_(True)
    E = rnd(V, K, gen=gen, scale=0.2).requires_grad_(True)
    tgt = torch.randint(2, V, (M,), generator=gen)
    tgt[::7] = 1
    nll = torch.zeros(1, dtype=torch.float32, device=dev())
    loss = ops.linear_cross_entropy(x, E, tgt.to(dev()), 1, eps, nll)
    loss.backward()
    xr, Er = x.detach().float().cpu().requires_grad_(True), E.detach().float().cpu().requires_grad_(True)
    logits_r = (xr @ Er.t()).to(torch.bfloat16).float()  # the kernel's logits are one bf16 scratch
    lr, nr, _ = om.label_smoothed_cross_entropy_sum(xr @ Er.t(), tgt, eps)
    lr.backward()
    assert abs(loss.item() - lr.item()) <= 3e-3 * abs(lr.item())
    assert abs(nll.item() - nr.item()) <= 3e-3 * abs(nr.item())
    assert rel_l2(x.grad, xr.grad) <= TOL16 and rel_l2(E.grad, Er.grad) <= TOL16


def test_grad_arena_matches_plain_backward():
    """GradArena (data-parallel gradient storage): the weight-gradient GEMMs write straight into the flat arena, leftovers are
    copied in, p.grad become arena views -- same numbers as the plain backward, packed q|k|v slots adjacent, an unused adaptor
    contributes zeros; gradients accumulate correctly over two backward passes of one step."""
    import ofasys_b200 as ob
    from ofasys_b200.distributed import GradArena
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import workloads

    cfg = ob.GeneralistModelConfig.default()
    cfg.dropout = cfg.attention_dropout = 0.0
    m = ob.GeneralistModel(cfg)
    for ad in ("text", "audio_fbank"):
        getattr(m.cfg.adaptor, ad).is_active = True
    torch.manual_seed(0)
    V = 600
    m.initialize(ob.Dictionary(n_dummy=V - 4))
    m = m.to(torch.bfloat16).to(dev()).train()
    gen = g()
    src = torch.randint(4, V, (3, 20), generator=gen).to(dev())
    prev, tgt = workloads._prev_target(gen, 3, 12, V)
    slots = [ob.Slot(ob.ModalityType.TEXT, True, src), ob.Slot(ob.ModalityType.TEXT, False, prev.to(dev()))]

    def run():
        m.forward_loss(slots, tgt.to(dev())).backward()

    run()  # (packs q|k|v storages)
    m.zero_grad(set_to_none=True)
    run()
    run()  # two task batches of one step: gradients accumulate
    ref = {k: None if p.grad is None else p.grad.clone() for k, p in m.named_parameters()}
    arena = GradArena(m.parameters(), bucket_bytes=1 << 20)
    try:
        assert len(arena.buckets) > 3
        arena.begin_step()
        run()
        arena.arm()
        run()
        arena.finish()
        torch.cuda.synchronize()
        lo, hi = arena.flat.data_ptr(), arena.flat.data_ptr() + arena.flat.numel() * 2
        n_direct = 0
        for k, p in m.named_parameters():
            assert p.grad is not None and lo <= p.grad.data_ptr() < hi, k  # every gradient lives in the arena
            if ref[k] is None:
                assert not p.grad.any(), k  # unused adaptor: zeros
            else:
                assert rel_l2(p.grad, ref[k]) <= 4e-3, (k, rel_l2(p.grad, ref[k]))  # (accumulation order of the two passes)
                n_direct += 1
        assert n_direct > 50
        a = m.encoder.layers[0].self_attn
        assert a.q_proj.weight.grad.data_ptr() + a.q_proj.weight.numel() * 2 == a.k_proj.weight.grad.data_ptr()  # packed slots
    finally:
        arena.close()


@pytest.mark.parametrize("tag,use_masks,use_range,dw", [("range", False, True, 0.0), ("masks", True, False, 0.0), ("both", True, True, 0.0), ("dropworst", False, True, 0.25)])
def test_constrained_label_smoothed_criterion(tag, use_masks, use_range, dw):
    """constraint_range / constraint masks / drop-worst of the label-smoothed criterion on the per-row CE
    kernels (ops.cross_entropy_rows + criterion.LabelSmoothedCrossEntropyCriterion) against the reference's own numbers
    (tests/golden/ls_ce_constraints.pt) and the oracle."""
    import os
    from ofasys_b200.criterion import LabelSmoothedCrossEntropyCriterion
    from oracle import oracle_model as om

    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ls_ce_constraints.pt"), weights_only=False)[tag]
    logits, target, masks, rng = om.make_constraint_case()
    x = logits.to(dev()).requires_grad_(True)

    class _M:  # the criterion only needs model(**net_input) -> (logits, extra) and get_targets
        def __call__(self, **kw):
            return x, {}

        def get_targets(self, sample, net_output):
            return sample["target"]

    sample = {"net_input": {}, "target": target.to(dev()), "ntokens": int((target != 1).sum()), "nsentences": target.shape[0]}
    if use_masks:
        sample["constraint_masks"] = masks.to(dev())
    crit = LabelSmoothedCrossEntropyCriterion(label_smoothing=0.1, drop_worst_ratio=dw, drop_worst_after=2,
                                              constraint_range=f"{rng[0]},{rng[1]}" if use_range else None)
    loss, sample_size, log = crit(_M(), sample, update_num=5)
    loss.backward()
    assert sample_size == fx["ntokens"]
    assert abs(float(loss) - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"]))
    assert abs(float(log["nll_loss"]) - float(fx["nll_loss"])) <= 2e-5 * abs(float(fx["nll_loss"]))
    assert rel_l2(x.grad.float(), fx["dlogits"]) <= 6e-3  # bf16 gradient storage
    disallowed = ~masks if use_masks else torch.zeros_like(masks)
    if use_range:
        disallowed = disallowed.clone()
        disallowed[..., 4:rng[0]] = True
        disallowed[..., rng[1]:] = True
    assert not x.grad.cpu()[disallowed].any()  # exact zeros outside the constraint set


def test_speech_to_text_criterion_ce_plus_ctc():
    """the composed ASR criterion (speech_to_text_loss.py:206-237): ce_weight * label-smoothed CE + ctc_weight *
    CTC on F.linear(encoder_out, E[dict_start:dict_end]) -- against the oracle model + the oracle's criterion pieces."""
    from ofasys_b200.criterion import SpeechToTextLossCriterion
    from oracle import cases
    from oracle import oracle_model as om
    from util import bf16_round_state_dict, build_product, load_golden as _lg, to_product_slots

    name = "audio_A"
    gold = _lg(name)
    sd = cases.synth_state_dict(gold["spec"], seed=0)
    sd_r = bf16_round_state_dict(sd)
    cfg = cases.oracle_cfg(name)
    slots, target = cases.make_inputs(name)
    m = build_product(name)
    m.load_state_dict(sd, strict=False)
    m = m.to(torch.bfloat16).to(dev()).train()
    gen = g()
    B = target.shape[0]
    d0, d1 = 100, 160  # phone range of the vocabulary; blank = first entry
    L = 9
    et = torch.randint(d0 + 1, d1, (B, L), generator=gen)
    et[:, -1] = 2  # eos
    et[1, 5:-1] = 1  # padding
    sample = {"net_input": {"slots": to_product_slots(slots, dev())}, "target": target.to(dev()), "ntokens": int((target != 1).sum()),
              "nsentences": B, "encoder_target": et.to(dev())}
    crit = SpeechToTextLossCriterion(d0, d1, blank_idx=0, ce_weight=0.7, ctc_weight=0.3, zero_infinity=True, label_smoothing=0.1)
    loss, sample_size, log = crit(m, sample)
    loss.backward()
    # oracle: the same composition from the oracle model
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd_r.items() if v.is_floating_point() and "running_" not in k}
    leaf["decoder.adaptor.embed_tokens.weight"] = leaf["encoder.adaptor.embed_tokens.weight"]
    full = dict(sd_r)
    full.update(leaf)
    logits, extra = om.model_forward(full, cfg, slots)
    ce, _, _ = om.constrained_criterion(logits, target, 0.1)
    enc = extra["encoder_out"]
    xc = torch.nn.functional.linear(enc["encoder_out"], full["decoder.adaptor.embed_tokens.weight"][d0:d1])  # T x B x C
    lens = (~enc["encoder_padding_mask"]).long().sum(-1)
    pm = (et != 1) & (et != 2)
    tl = pm.sum(-1)
    lp = torch.log_softmax(xc.float(), -1)
    ctc = torch.nn.functional.ctc_loss(lp, (et - d0).masked_select(pm), lens, tl, blank=0, reduction="sum", zero_infinity=True)
    ref = 0.7 * ce + 0.3 * ctc
    ref.backward()
    assert abs(float(log["ctc_loss"]) - float(ctc)) <= 5e-3 * abs(float(ctc)), (float(log["ctc_loss"]), float(ctc))
    assert abs(float(loss) - float(ref)) <= 3e-3 * abs(float(ref)), (float(loss), float(ref))
    gE = dict(m.named_parameters())["encoder.adaptor.embed_tokens.weight"].grad.float().cpu()
    assert rel_l2(gE, leaf["encoder.adaptor.embed_tokens.weight"].grad) <= 3e-2
    k = "encoder.layers.1.fc1.weight"
    assert rel_l2(dict(m.named_parameters())[k].grad.float().cpu(), leaf[k].grad) <= 3e-2
