"""Per-kernel timing of the LayerNorm family at the shapes of the benchmark step (B=32), with achieved HBM GB/s
against the algorithmic bytes.  Development tool; not part of the product path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_it(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    out = []
    for tag, rows in (("enc32", 8480), ("enc64", 16960), ("dec64", 4096)):
        # GELU + ffn_layernorm
        cols = 3072
        x = torch.randn(rows, cols, device=dev).bfloat16().requires_grad_(True)
        w = (torch.rand(cols, device=dev) + 0.5).bfloat16().requires_grad_(True)
        b = torch.randn(cols, device=dev).bfloat16().requires_grad_(True)
        y = ops.layer_norm(x, w, b, gelu=True)
        dy = torch.randn_like(y)
        t = time_it(lambda: ops.layer_norm(x.detach(), w.detach(), b.detach(), gelu=True))
        out.append((f"{tag}.gelu_ln.fwd", rows, cols, t, rows * cols * 4))
        t = time_it(lambda: torch.autograd.grad(y, (x, w, b), dy, retain_graph=True))
        out.append((f"{tag}.gelu_ln.bwd(+reduce)", rows, cols, t, rows * cols * 6))
        # junction LN -> +res -> LN
        cols = 768
        a = torch.randn(rows, cols, device=dev).bfloat16().requires_grad_(True)
        xr = torch.randn(rows, cols, device=dev).requires_grad_(True)
        ws = [(torch.rand(cols, device=dev) + 0.5).bfloat16().requires_grad_(True) for _ in range(4)]
        for has1 in (True, False):
            w1, b1 = (ws[0], ws[1]) if has1 else (None, None)
            xn, yy = ops.ln_res_ln(a, xr, w1, b1, ws[2], ws[3])
            g1, g2 = torch.randn_like(xn), torch.randn_like(yy)
            t = time_it(lambda: ops.ln_res_ln(a.detach(), xr.detach(), None if w1 is None else w1.detach(), None if b1 is None else b1.detach(), ws[2].detach(), ws[3].detach()))
            out.append((f"{tag}.ln_res_ln<{has1}>.fwd", rows, cols, t, rows * cols * 12))
            ins = (a, xr, ws[2], ws[3]) + ((w1, b1) if has1 else ())
            t = time_it(lambda: torch.autograd.grad((xn, yy), ins, (g1, g2), retain_graph=True))
            out.append((f"{tag}.ln_res_ln<{has1}>.bwd(+reduce)", rows, cols, t, rows * cols * (18 if has1 else 16)))
    recs = []
    for name, rows, cols, t, byts in out:
        print(f"{name:34s} rows={rows:6d} cols={cols:5d} {t * 1e6:8.1f} us  {byts / t / 1e9:8.1f} GB/s (algorithmic bytes {byts / 1e6:.1f} MB)")
        recs.append({"name": name, "rows": rows, "cols": cols, "us": t * 1e6, "gbs": byts / t / 1e9, "bytes": byts})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(recs, open("gpurun_out/ln_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
