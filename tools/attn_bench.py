"""Attention fwd / bwd kernel time on the shapes of the BASELINE workloads: each call sequence is captured into a CUDA graph and
replayed (CUDA events around 20 replays), so the numbers are device time, not host launch overhead."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200 import ops
dev = torch.device("cuda:0")
DENSE = os.environ.get("ATTN_DENSE", "1") == "1"

def graph_time(fn, reps=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

BB = int(os.environ.get("ATTN_B", "32"))
for name, Tq, Tk, causal, mode, modeA in [("enc self 265", 265, 265, False, "self", False), ("dec self 64 causal", 64, 64, True, "self", False),
                                         ("dec cross 64x265", 64, 265, False, "cross", False), ("enc self 204 modeA", 204, 204, False, "self", True),
                                         ("enc self 260 modeA", 260, 260, False, "self", True), ("dec self 128 causal modeA", 128, 128, True, "self", True),
                                         ("dec cross 128x260 modeA", 128, 260, False, "cross", True), ("enc self 1040", 1040, 1040, False, "self", False),
                                         ("enc self 1040 modeA", 1040, 1040, False, "self", True), ("enc self 3144 modeA", 3144, 3144, False, "self", True)]:
    b = BB if Tq < 1000 else (8 if Tq < 2000 else 2)
    H = 12 if Tq < 1000 else 16
    d = H * 64
    qs = torch.randn(b, Tq, 3 * d if mode == "self" else d, device=dev).bfloat16().requires_grad_(True)
    kv = None if mode == "self" else torch.randn(b, Tk, 2 * d, device=dev).bfloat16().requires_grad_(True)
    kpm = torch.zeros(b, Tk, dtype=torch.bool, device=dev)
    pq = pk = tab = idx = None
    if modeA:
        pq = torch.randn(1, Tq, d, device=dev).bfloat16().requires_grad_(True); pk = torch.randn(1, Tk, d, device=dev).bfloat16().requires_grad_(True)
        if mode == "self":
            tab = torch.randn(7000, H, device=dev).bfloat16().requires_grad_(True)
            idx = torch.randint(0, 7000, (Tq, Tk), device=dev, dtype=torch.int32)
    do = torch.randn(b, Tq, d, device=dev).bfloat16()
    ab = ops.abs_pos(pq, pk, H).detach() if (modeA and DENSE) else None  # per forward, shared by the layers: not part of a layer's time
    def bias():
        return ops.PositionBias(pq, pk, idx, tab, abs=None if ab is None else ab.requires_grad_(True)) if modeA else None
    def fwd():
        with torch.no_grad(): ops.attention(qs, kv, H, 0.125, bias(), kpm, causal)
    def fb():
        o = ops.attention(qs, kv, H, 0.125, bias(), kpm, causal); o.backward(do)
        qs.grad = None
    tf, tfb = graph_time(fwd), graph_time(fb)
    fl = 4.0 * b * H * Tq * Tk * 64 * (0.5 if causal else 1.0)
    print(f"{name:26s} B={b:3d} fwd {tf:8.1f} us ({fl / tf / 1e6:7.1f} TF/s)   bwd {tfb - tf:8.1f} us ({2.5 * fl / (tfb - tf) / 1e6:7.1f} TF/s alg)", flush=True)
