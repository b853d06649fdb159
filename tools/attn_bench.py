"""Attention fwd / bwd timing on the shapes of the benchmark step (B=32, H=12)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200 import ops
dev = torch.device("cuda:0")

def ev(fn, iters=15):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

B, H = 32, 12
d = H * 64
for name, Tq, Tk, causal, mode, modeA in [("enc self 265", 265, 265, False, "self", False), ("dec self 64 causal", 64, 64, True, "self", False),
                                         ("dec cross 64x265", 64, 265, False, "cross", False), ("enc self 204 modeA", 204, 204, False, "self", True),
                                         ("enc self 1040", 1040, 1040, False, "self", False)]:
    b = B if Tq < 1000 else 8
    qs = torch.randn(b, Tq, 3 * d if mode == "self" else d, device=dev).bfloat16().requires_grad_(True)
    kv = None if mode == "self" else torch.randn(b, Tk, 2 * d, device=dev).bfloat16().requires_grad_(True)
    kpm = torch.zeros(b, Tk, dtype=torch.bool, device=dev)
    bias = None
    if modeA:
        pq = torch.randn(1, Tq, d, device=dev).bfloat16().requires_grad_(True); pk = torch.randn(1, Tk, d, device=dev).bfloat16().requires_grad_(True)
        tab = torch.randn(7000, H, device=dev).bfloat16().requires_grad_(True)
        idx = torch.randint(0, 7000, (Tq, Tk), device=dev, dtype=torch.int32)
        bias = ops.PositionBias(pq, pk, idx, tab)
    o = ops.attention(qs, kv, H, 0.125, bias, kpm, causal)
    do = torch.randn_like(o)
    tf = ev(lambda: ops.attention(qs, kv, H, 0.125, bias, kpm, causal))
    def fb():
        o = ops.attention(qs, kv, H, 0.125, bias, kpm, causal); o.backward(do)
    tfb = ev(fb)
    fl = 4.0 * b * H * Tq * Tk * 64 * (0.5 if causal else 1.0)
    print(f"{name:22s} fwd {tf:7.1f} us ({fl / tf / 1e6:6.1f} TF/s)   bwd {tfb - tf:7.1f} us ({2.5 * fl / (tfb - tf) / 1e6:6.1f} TF/s alg)")
