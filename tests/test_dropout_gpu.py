"""GPU: dropout / drop-path (ofasys/module/dropout.py:14-25, droppath.py:13-63).

The reference draws masks from torch's RNG, the CUDA path from a counter-based hash (no mask tensor in HBM), so parity
is checked with the SAME masks: ops.dropout_mask() materialises the multipliers a descriptor applies and the
oracle multiplies them in at the reference's Dropout / DropPath call sites (oracle_model.DROP_HOOK).  Tolerances are
those of tests/test_model_gpu.py (bf16 GEMM operands).  Mask statistics are tested separately.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import cases
from oracle import oracle_model as om
from util import bf16_round_state_dict, build_product, load_golden, rel_l2, to_product_slots

pytestmark = pytest.mark.gpu


def _ops():
    from ofasys_b200 import ops
    return ops


def test_mask_statistics_and_determinism():
    ops = _ops()
    st = ops.DropoutState("cuda:0", seed=1234)
    st.next_step()
    d1 = st.spec(0.1)
    d2 = st.spec(0.1)
    m1 = ops.dropout_mask(d1, 4096, 768)
    m1b = ops.dropout_mask(d1, 4096, 768)
    m2 = ops.dropout_mask(d2, 4096, 768)
    assert torch.equal(m1, m1b)  # same descriptor, same step -> same mask (what backward relies on)
    assert not torch.equal(m1, m2)  # another call site
    vals = torch.unique(m1)
    assert vals.numel() == 2 and vals[0].item() == 0.0 and abs(vals[1].item() - 1 / 0.9) < 1e-6
    frac = (m1 == 0).float().mean().item()
    assert abs(frac - 0.1) < 2e-3, frac  # 3.1M draws: sigma = 1.7e-4
    # per-column and per-row rates are flat (no stripe structure from the 8-column counter blocks)
    assert ((m1 == 0).float().mean(0) - 0.1).abs().max().item() < 0.03
    assert ((m1 == 0).float().mean(1) - 0.1).abs().max().item() < 0.06
    # independence of neighbours inside one 8-column block
    z = (m1 == 0).float()
    corr = ((z[:, :-1] - 0.1) * (z[:, 1:] - 0.1)).mean().item() / 0.09
    assert abs(corr) < 5e-3, corr
    st.next_step()  # a new step -> new mask for the same site
    assert not torch.equal(m1, ops.dropout_mask(d1, 4096, 768))


def test_drop_path_zeroes_whole_samples():
    ops = _ops()
    st = ops.DropoutState("cuda:0", seed=7)
    st.next_step()
    B, T, C = 512, 16, 64
    d = st.spec(0.0, drop_path=0.25, rows_per_sample=T)
    m = ops.dropout_mask(d, B * T, C).view(B, T * C)
    per_sample = m[:, 0]
    assert torch.equal(m, per_sample[:, None].expand_as(m))  # one value per sample
    kept = per_sample != 0
    assert torch.allclose(per_sample[kept], torch.full_like(per_sample[kept], 1 / 0.75))
    assert abs((~kept).float().mean().item() - 0.25) < 0.06
    # combined with element dropout: kept elements carry both scales
    d2 = st.spec(0.5, drop_path=0.25, rows_per_sample=T)
    m2 = ops.dropout_mask(d2, B * T, C)
    vals = torch.unique(m2)
    assert vals.numel() == 2 and vals[0].item() == 0.0 and abs(vals[1].item() - 1 / 0.5 / 0.75) < 1e-5


def test_standalone_dropout_fwd_bwd():
    ops = _ops()
    st = ops.DropoutState("cuda:0", seed=3)
    st.next_step()
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(37, 5, 256, device="cuda", dtype=dt, requires_grad=True)
        d = st.spec(0.3)
        y = ops.dropout(x, d)
        m = ops.dropout_mask(d, 37 * 5, 256).view(37, 5, 256)
        assert torch.equal(y.float(), (x.detach().float() * m).to(dt).float())
        g = torch.randn_like(y)
        y.backward(g)
        assert torch.equal(x.grad.float(), (g.float() * m).to(dt).float())
    assert ops.dropout(x, None) is x


@pytest.mark.parametrize("has_ln1", [True, False])
@pytest.mark.parametrize("cols,rows,T", [(768, 530, 53), (256, 96, 12), (1024, 64, 8)])
def test_ln_res_ln_dropout(has_ln1, cols, rows, T):
    ops = _ops()
    torch.manual_seed(0)
    st = ops.DropoutState("cuda:0", seed=11)
    st.next_step()
    B = rows // T
    a = torch.randn(B, T, cols, device="cuda").bfloat16().requires_grad_(True)
    x = torch.randn(B, T, cols, device="cuda").requires_grad_(True)
    w1 = (torch.rand(cols, device="cuda") + 0.5).bfloat16().requires_grad_(True) if has_ln1 else None
    b1 = (0.1 * torch.randn(cols, device="cuda")).bfloat16().requires_grad_(True) if has_ln1 else None
    w2 = (torch.rand(cols, device="cuda") + 0.5).bfloat16().requires_grad_(True)
    b2 = (0.1 * torch.randn(cols, device="cuda")).bfloat16().requires_grad_(True)
    d = st.spec(0.1, drop_path=0.2, rows_per_sample=T)
    xn, y = ops.ln_res_ln(a, x, w1, b1, w2, b2, 1e-5, drop=d)
    m = ops.dropout_mask(d, rows, cols).view(B, T, cols)
    gx, gy = torch.randn_like(xn), torch.randn_like(y)
    (xn * gx).sum().add((y.float() * gy.float()).sum()).backward()

    def ref():
        a_, x_ = a.detach().float().requires_grad_(True), x.detach().requires_grad_(True)
        ps = [None if p is None else p.detach().float().requires_grad_(True) for p in (w1, b1, w2, b2)]
        br = F.layer_norm(a_, (cols,), ps[0], ps[1], 1e-5) if has_ln1 else a_
        xn_ = x_ + m * br
        y_ = F.layer_norm(xn_, (cols,), ps[2], ps[3], 1e-5)
        (xn_ * gx).sum().add((y_ * gy.float()).sum()).backward()
        return xn_, y_, a_.grad, x_.grad, [None if p is None else p.grad for p in ps]

    xn_r, y_r, da_r, dx_r, dps = ref()
    assert rel_l2(xn, xn_r) < 1e-5
    assert rel_l2(y.float(), y_r) < 4e-3
    assert rel_l2(x.grad, dx_r) < 4e-3
    assert rel_l2(a.grad.float(), da_r) < 8e-3
    for p, gr in zip((w1, b1, w2, b2), dps):
        if p is not None:
            assert rel_l2(p.grad.float(), gr) < 1.5e-2


@pytest.mark.parametrize("cols,rows", [(3072, 530), (1024, 96), (4096, 40)])
def test_gelu_ln_activation_dropout(cols, rows):
    ops = _ops()
    torch.manual_seed(1)
    st = ops.DropoutState("cuda:0", seed=5)
    st.next_step()
    x = torch.randn(rows, cols, device="cuda").bfloat16().requires_grad_(True)
    w = (torch.rand(cols, device="cuda") + 0.5).bfloat16().requires_grad_(True)
    b = (0.1 * torch.randn(cols, device="cuda")).bfloat16().requires_grad_(True)
    d = st.spec(0.15)
    y = ops.layer_norm(x, w, b, 1e-5, gelu=True, drop=d)
    m = ops.dropout_mask(d, rows, cols)
    gy = torch.randn_like(y)
    (y.float() * gy.float()).sum().backward()
    x_, w_, b_ = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    y_r = F.layer_norm(F.gelu(x_) * m, (cols,), w_, b_, 1e-5)
    (y_r * gy.float()).sum().backward()
    assert rel_l2(y.float(), y_r) < 4e-3
    assert rel_l2(x.grad.float(), x_.grad) < 8e-3
    assert rel_l2(w.grad.float(), w_.grad) < 1.5e-2 and rel_l2(b.grad.float(), b_.grad) < 1.5e-2


def test_embed_ln_dropout():
    ops = _ops()
    torch.manual_seed(2)
    st = ops.DropoutState("cuda:0", seed=9)
    st.next_step()
    B, T, d_, V = 3, 17, 256, 97
    E = torch.randn(V, d_, device="cuda").bfloat16().requires_grad_(True)
    pos = torch.randn(32, d_, device="cuda").bfloat16().requires_grad_(True)
    g = (torch.rand(d_, device="cuda") + 0.5).bfloat16().requires_grad_(True)
    be = (0.1 * torch.randn(d_, device="cuda")).bfloat16().requires_grad_(True)
    tok = torch.randint(2, V, (B, T), device="cuda")
    d = st.spec(0.2)
    out = ops.embed_ln(g, be, tokens=tok, E=E, pos=pos, drop=d, padding_idx=1)
    m = ops.dropout_mask(d, B * T, d_).view(B, T, d_)
    go = torch.randn_like(out)
    (out * go).sum().backward()
    E_, pos_, g_, be_ = (t.detach().float().requires_grad_(True) for t in (E, pos, g, be))
    ref = F.layer_norm(F.embedding(tok, E_) + pos_[:T], (d_,), g_, be_, 1e-5) * m
    (ref * go).sum().backward()
    assert rel_l2(out, ref) < 1e-5
    assert rel_l2(E.grad.float(), E_.grad) < 8e-3
    assert rel_l2(pos.grad.float(), pos_.grad) < 8e-3
    assert rel_l2(g.grad.float(), g_.grad) < 1.5e-2


def _attn_ref(q, k, v, kpm, causal, scale, H, mask):
    """multihead_attention.py:308-338 in fp32 with the dropout multipliers applied to the probabilities (:335)."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    sp = lambda t: t.view(t.shape[0], t.shape[1], H, 64).transpose(1, 2)
    s = torch.matmul(sp(q), sp(k).transpose(2, 3)) * scale
    if causal:
        s = s + torch.triu(torch.full((Tq, Tk), float("-inf"), device=s.device), 1)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1) * mask
    return torch.matmul(p, sp(v)).transpose(1, 2).reshape(B, Tq, d)


@pytest.mark.parametrize("mode,Tq,Tk,causal,use_kpm", [("self", 130, 130, False, True), ("self", 64, 64, True, False),
                                                      ("cross", 24, 265, False, True), ("self", 257, 257, True, True)])
def test_attention_dropout(mode, Tq, Tk, causal, use_kpm):
    ops = _ops()
    torch.manual_seed(4)
    st = ops.DropoutState("cuda:0", seed=17)
    st.next_step()
    B, H = 2, 3
    d = H * 64
    scale = 128 ** -0.5
    if mode == "self":
        qkv = torch.randn(B, Tq, 3 * d, device="cuda").bfloat16().requires_grad_(True)
        kv = None
    else:
        qkv = torch.randn(B, Tq, d, device="cuda").bfloat16().requires_grad_(True)
        kv = torch.randn(B, Tk, 2 * d, device="cuda").bfloat16().requires_grad_(True)
    kpm = None
    if use_kpm:
        kpm = torch.zeros(B, Tk, dtype=torch.bool, device="cuda")
        kpm[1, Tk - Tk // 4:] = True
        kpm[0, 3] = True
    drop = st.spec(0.25)
    o = ops.attention(qkv, kv, H, scale, None, kpm, causal, drop=drop)
    do = torch.randn_like(o)
    o.backward(do)
    mask = ops.attention_dropout_mask(drop, B, H, Tq, Tk)
    frac = (mask == 0).float().mean().item()
    assert abs(frac - 0.25) < 0.02, frac
    # no structure along either axis (the hash is of i * Tk + j)
    assert ((mask == 0).float().mean(dim=(0, 1, 2)) - 0.25).abs().max().item() < 0.12
    assert ((mask == 0).float().mean(dim=(0, 1, 3)) - 0.25).abs().max().item() < 0.12
    refs = [t.detach().float().requires_grad_(True) for t in (qkv, kv) if t is not None]
    if mode == "self":
        q_, k_, v_ = refs[0][..., :d], refs[0][..., d:2 * d], refs[0][..., 2 * d:]
    else:
        q_, k_, v_ = refs[0], refs[1][..., :d], refs[1][..., d:]
    orf = _attn_ref(q_, k_, v_, kpm, causal, scale, H, mask)
    orf.backward(do.float())
    assert rel_l2(o, orf) <= 1e-2
    for t, r in zip([t for t in (qkv, kv) if t is not None], refs):
        assert rel_l2(t.grad, r.grad) <= 2e-2, tuple(t.shape)


def test_graph_replay_draws_fresh_masks():
    """state lives in device memory and next_step() is device work: every replay of a captured step sees a new mask."""
    ops = _ops()
    st = ops.DropoutState("cuda:0", seed=21)
    x = torch.ones(64, 256, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        st.next_step()
        d = st.spec(0.5)
        y = ops.dropout(x, d)
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        st.next_step()
        d = st.spec(0.5)
        y = ops.dropout(x, d)
    outs = []
    for _ in range(3):
        gr.replay()
        torch.cuda.synchronize()
        outs.append(y.clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    for o in outs:
        assert abs((o == 0).float().mean().item() - 0.5) < 0.03


class _Recorder:
    """Records the descriptors a product forward creates and replays them, in order, as the oracle's masks."""

    def __init__(self, ops, active, heads=1):
        self.ops, self.active, self.specs, self.i, self.heads = ops, active, [], 0, heads
        self._orig = ops.DropoutState.spec

    def __enter__(self):
        rec = self

        def spec(self_, *a, **k):
            d = rec._orig(self_, *a, **k)
            if d is not None:
                rec.specs.append(d)
            return d

        self.ops.DropoutState.spec = spec
        return self

    def __exit__(self, *exc):
        self.ops.DropoutState.spec = self._orig

    def hook(self, kind, x):
        if not self.active.get(kind, False):
            return torch.ones((), dtype=x.dtype)
        d = self.specs[self.i]
        self.i += 1
        if kind == "embed":  # B x T x C
            B, T, C = x.shape
            return self.ops.dropout_mask(d, B * T, C).view(B, T, C).cpu().to(x.dtype)
        if kind in ("branch", "act"):  # T x B x C in the reference layout; kernel rows are b * T + t
            T, B, C = x.shape
            return self.ops.dropout_mask(d, B * T, C).view(B, T, C).transpose(0, 1).cpu().to(x.dtype)
        if kind == "attn_probs":  # B*H x T x S
            BH, T, S = x.shape
            return self.ops.attention_dropout_mask(d, BH // self.heads, self.heads, T, S).view(BH, T, S).cpu().to(x.dtype)
        raise AssertionError(kind)


@pytest.mark.parametrize("name", ["text_A", "patch_B"])
def test_model_parity_with_dropout(name):
    """Whole model in training mode with dropout 0.1 (the reference default, config/default_model.yaml:15), activation
    dropout 0.1 and drop-path 0.2: logits, loss and every gradient against the oracle using the same masks."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = load_golden(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    sd_r = bf16_round_state_dict(sd)
    cfg = cases.oracle_cfg(name)
    slots, target = cases.make_inputs(name)

    m = build_product(name)
    m.load_state_dict(sd, strict=False)
    m = m.to(torch.bfloat16).to(dev).train()
    p_res, p_act, p_path, p_attn = 0.1, 0.1, 0.2, 0.1
    L = len(m.encoder.layers)
    for i, layer in enumerate(list(m.encoder.layers) + list(m.decoder.layers)):
        layer.dropout_p, layer.activation_dropout_p = p_res, p_act
        layer.drop_path_rate = p_path * (i % L) / max(L - 1, 1)  # transformer.py:58,249 linspace(0, rate, L) schedule
        for att in (layer.self_attn, getattr(layer, "encoder_attn", None)):
            if att is not None:
                att.dropout_p = p_attn
    for mod in m.modules():
        if hasattr(mod, "hook") and hasattr(mod, "dropout_p"):
            mod.dropout_p = p_res
    m.encoder._uses_dropout = True
    ops.dropout_state(dev).reseed(2024)
    pslots = to_product_slots(slots, dev)
    tgt = target.to(dev)
    active = {"embed": True, "branch": True, "act": True, "attn_probs": True}
    H = cfg.heads

    # (1) logits through the reference-facing API
    with _Recorder(ops, active, H) as rec:
        logits, _ = m(pslots)
        torch.cuda.synchronize()
        om.DROP_HOOK = rec.hook
        try:
            with torch.no_grad():
                logits_ref, _ = om.model_forward(sd_r, cfg, slots)
        finally:
            om.DROP_HOOK = None
        assert rec.i == len(rec.specs), (rec.i, len(rec.specs))  # every descriptor matched a reference call site
    e_logits = rel_l2(logits.float(), logits_ref)

    # the masks matter: the eval-mode logits are far away
    with torch.no_grad():
        logits_eval, _ = om.model_forward(sd_r, cfg, slots)
    assert rel_l2(logits_ref, logits_eval) > 0.05

    # (2) measured path: fused projection + criterion + backward (a new step -> new masks)
    with _Recorder(ops, active, H) as rec:
        m.zero_grad(set_to_none=True)
        loss = m.forward_loss(pslots, tgt)
        loss.backward()
        torch.cuda.synchronize()
        om.DROP_HOOK = rec.hook
        try:
            loss_ref, _, grads_ref = om.loss_and_grads(sd_r, cfg, slots, target)
        finally:
            om.DROP_HOOK = None
        assert rec.i == len(rec.specs)
    e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    num = den = 0.0
    worst = (None, 0.0)
    for k, p in m.named_parameters():
        gr = grads_ref[k].double()
        gp = (torch.zeros_like(p) if p.grad is None else p.grad).double().cpu()
        num += (gp - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    for k, p in m.named_parameters():
        gr = grads_ref[k].double()
        if gr.norm() > 1e-3 * den ** 0.5:
            e = ((p.grad.double().cpu() - gr).norm() / gr.norm()).item()
            if e > worst[1]:
                worst = (k, e)
    e_grad = (num / max(den, 1e-30)) ** 0.5
    assert e_logits <= 6e-3, e_logits
    assert e_loss <= 2e-3, (loss.item(), loss_ref.item())
    assert e_grad <= 3e-2, e_grad
    assert worst[1] <= 8e-2, worst

    # eval mode ignores dropout entirely (dropout.py:21-25)
    m.eval()
    with _Recorder(ops, active) as rec:
        with torch.no_grad():
            le, _ = m(pslots)
        assert not rec.specs
    assert rel_l2(le.float(), logits_eval) <= 6e-3
