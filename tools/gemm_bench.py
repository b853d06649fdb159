"""Per-shape timing of ofab_gemm_bf16 on the GEMM shapes of the benchmark step (B=32), against cuBLAS
(torch.matmul) as the measured ceiling.  Development tool; not part of the product path."""
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


def bench(name, M, N, K, a_mn, b_mn, out_dtype=torch.bfloat16):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    Am = A.t().contiguous() if a_mn else A
    Bm = B.t().contiguous() if b_mn else B
    Np = (N + 7) // 8 * 8
    out = torch.empty(M, Np, dtype=out_dtype, device=dev)
    t = time_it(lambda: (ops.gemm_splitk if (a_mn and b_mn and out_dtype == torch.bfloat16) else ops.gemm)(M, N, K, Am, Am.stride(0), a_mn, Bm, Bm.stride(0), b_mn, out, Np))
    tc = time_it(lambda: torch.matmul(A, B.t()))
    fl = 2.0 * M * N * K
    return {"name": name, "M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "us": t * 1e6, "tflops": fl / t / 1e12,
            "cublas_us": tc * 1e6, "cublas_tflops": fl / tc / 1e12}


def main():
    rows = []
    BS = int(os.environ.get("GEMM_B", "32"))  # per-GPU batch: rows = BS x 265 (encoder) / BS x 64 (decoder)
    ME, MD = BS * 265, BS * 64
    for tag, M in (("enc", ME), ("dec", MD)):
        for lname, N, K in (("qkv", 2304, 768), ("out", 768, 768), ("fc1", 3072, 768), ("fc2", 768, 3072)):
            rows.append(bench(f"{tag}.{lname}.fwd", M, N, K, 0, 0))
            rows.append(bench(f"{tag}.{lname}.dgrad", M, K, N, 0, 1))
            rows.append(bench(f"{tag}.{lname}.wgrad", N, K, M, 1, 1))
    rows.append(bench("dec.crosskv.fwd", ME, 1536, 768, 0, 0))
    rows.append(bench("dec.crosskv.dgrad", ME, 768, 1536, 0, 1))
    rows.append(bench("dec.crosskv.wgrad", 1536, 768, ME, 1, 1))
    rows.append(bench("logits.fwd", MD, 50264, 768, 0, 0))
    rows.append(bench("logits.dgrad", MD, 768, 50264, 0, 1))
    rows.append(bench("logits.wgrad", 50264, 768, MD, 1, 1))
    tot = sum(r["us"] for r in rows if not r["name"].startswith("logits") and "crosskv" not in r["name"]) * 12
    tot += sum(r["us"] for r in rows if "crosskv" in r["name"]) * 12 + sum(r["us"] for r in rows if r["name"].startswith("logits"))
    for r in rows:
        print(f"{r['name']:22s} M={r['M']:6d} N={r['N']:6d} K={r['K']:6d}  {r['us']:8.1f} us {r['tflops']:7.1f} TF/s | cuBLAS {r['cublas_us']:8.1f} us {r['cublas_tflops']:7.1f} TF/s")
    print(f"estimated GEMM time per step (12 layers each): {tot / 1e3:.2f} ms")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
