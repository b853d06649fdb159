"""Bring-up check of the tcgen05 attention kernels against a plain fp32 reference, piece by piece (o, lse, dq, dk, dv,
bias gradients), on shapes that exercise one block, tails, several blocks, masks and the dense bias."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ofasys_b200 import ops
from test_ops_gpu import attn_ref

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def run(mode, Tq, Tk, variant, B=2, H=2):
    g = torch.Generator().manual_seed(1)
    d = H * 64
    scale = 128 ** -0.5
    R = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).bfloat16().to(dev)
    if mode == "self":
        qkv = R(B, Tq, 3 * d).requires_grad_(True); kv = None
    else:
        qkv = R(B, Tq, d).requires_grad_(True); kv = R(B, Tk, 2 * d).requires_grad_(True)
    pq = pk = table = idx = kpm = None
    if "pos" in variant:
        pq = R(1, Tq, d).requires_grad_(True); pk = R(1, Tk, d).requires_grad_(True)
    if "rel" in variant:
        table = R(37, H, sc=0.5).requires_grad_(True)
        idx = torch.randint(-1, 37, (Tq, Tk), generator=g).to(torch.int32).to(dev)
    if "kpm" in variant:
        kpm = torch.zeros(B, Tk, dtype=torch.bool); kpm[1, Tk - max(1, Tk // 4):] = True
        if Tk > 8: kpm[0, 3] = True
        kpm = kpm.to(dev)
    causal = "causal" in variant
    do = R(B, Tq, d)
    bias = ops.PositionBias(pq, pk, idx, table, abs=ops.abs_pos(pq, pk, H) if pq is not None else None)
    o = ops.attention(qkv, kv, H, scale, bias, kpm, causal)
    o.backward(do)
    torch.cuda.synchronize()
    leaves = [t for t in (qkv, kv, pq, pk, table) if t is not None]
    refs = {id(t): t.detach().float().requires_grad_(True) for t in leaves}
    Rf = lambda t: None if t is None else refs[id(t)]
    if mode == "self":
        q_, k_, v_ = Rf(qkv)[..., :d], Rf(qkv)[..., d:2 * d], Rf(qkv)[..., 2 * d:]
    else:
        q_, k_, v_ = Rf(qkv), Rf(kv)[..., :d], Rf(kv)[..., d:]
    pq_ = None if pq is None else Rf(pq).expand(B, -1, -1)
    pk_ = None if pk is None else Rf(pk).expand(B, -1, -1)
    orf = attn_ref(q_, k_, v_, pq_, pk_, Rf(table), idx, kpm, causal, scale, H)
    orf.backward(do.float())
    out = [f"o {rel(o, orf):.2e}"]
    if mode == "self":
        gq, gr = qkv.grad, Rf(qkv).grad
        out += [f"dq {rel(gq[..., :d], gr[..., :d]):.2e}", f"dk {rel(gq[..., d:2*d], gr[..., d:2*d]):.2e}", f"dv {rel(gq[..., 2*d:], gr[..., 2*d:]):.2e}"]
    else:
        out += [f"dq {rel(qkv.grad, Rf(qkv).grad):.2e}", f"dk {rel(kv.grad[..., :d], Rf(kv).grad[..., :d]):.2e}", f"dv {rel(kv.grad[..., d:], Rf(kv).grad[..., d:]):.2e}"]
    for nm, t in (("dpq", pq), ("dpk", pk), ("dtab", table)):
        if t is not None:
            out.append(f"{nm} {rel(t.grad, Rf(t).grad):.2e}")
    bad = any(float(x.split()[1]) > 2e-2 or x.split()[1] == "nan" for x in out)
    print(f"{'BAD ' if bad else 'ok  '}{mode:5s} {Tq:4d}x{Tk:4d} {variant:20s} " + "  ".join(out), flush=True)

cases = [("self", 128, 128, "plain"), ("self", 64, 64, "plain"), ("self", 24, 24, "plain"), ("cross", 16, 265, "plain"), ("self", 130, 130, "plain"),
         ("self", 257, 257, "plain"), ("self", 700, 700, "plain"), ("self", 64, 64, "causal"), ("self", 257, 257, "causal"), ("self", 130, 130, "kpm"),
         ("cross", 64, 265, "kpm"), ("self", 64, 64, "pos"), ("self", 130, 130, "pos_rel"), ("self", 257, 257, "pos_rel_kpm_causal"),
         ("cross", 16, 265, "pos"), ("self", 520, 520, "pos_rel_kpm")]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for c in cases:
    try:
        run(*c)
    except Exception as ex:
        print(f"EXC  {c}: {type(ex).__name__}: {str(ex)[:300]}", flush=True)
        break
