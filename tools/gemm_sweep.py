"""Sweep (cta_group, BN) per GEMM shape of the benchmark step in ONE process (env overrides are read per call)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def time_it(fn, iters=12):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2] * 1e3

shapes = []
for tag, M in (("enc", 8480), ("dec", 2048)):
    for l, N, K in (("qkv", 2304, 768), ("out", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)):
        shapes += [(f"{tag}.{l}.fwd", M, N, K, 0, 0), (f"{tag}.{l}.dgrad", M, K, N, 0, 1), (f"{tag}.{l}.wgrad", N, K, M, 1, 1)]
shapes += [("crosskv.fwd", 8480, 1536, 768, 0, 0), ("crosskv.dgrad", 8480, 768, 1536, 0, 1), ("crosskv.wgrad", 1536, 768, 8480, 1, 1),
           ("logits.fwd", 2048, 50264, 768, 0, 0), ("logits.dgrad", 2048, 768, 50264, 0, 1), ("logits.wgrad", 50264, 768, 2048, 1, 1)]
rows = []
for name, M, N, K, a_mn, b_mn in shapes:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    Am = A.t().contiguous() if a_mn else A; Bm = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res = {}
    for cg in (1, 2):
        for bn in (128, 256):
            os.environ["OFAB_GEMM_CG"] = str(cg); os.environ["OFAB_GEMM_BN"] = str(bn)
            res[f"cg{cg}bn{bn}"] = time_it(lambda: ops.gemm(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out, N))
    os.environ.pop("OFAB_GEMM_CG"); os.environ.pop("OFAB_GEMM_BN")
    res["auto"] = time_it(lambda: ops.gemm(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out, N))
    best = min((v, k) for k, v in res.items() if k != "auto")
    print(f"{name:18s} M={M:6d} N={N:6d} K={K:6d} " + " ".join(f"{k}={v:6.1f}" for k, v in res.items()) + f"  best={best[1]}", flush=True)
    rows.append(dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, **res))
json.dump(rows, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
