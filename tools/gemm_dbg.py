"""Which pipeline stage bounds the GEMM?  Times one shape with parts of the kernel disabled (OFAB_GEMM_DBG)."""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ofasys_b200 import ops
    dev = torch.device("cuda:0")
    M, N, K = [int(x) for x in sys.argv[2:5]]
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(M, N, K, A, K, 0, B, K, 0, out, N)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    print(f"{ts[len(ts)//2]*1e3:8.1f} us  {2.0*M*N*K/ts[len(ts)//2]/1e9:7.1f} TF/s")
else:
    cases = [((8480, 768, 3072), 2, 128), ((8480, 3072, 768), 2, 256)] if "--quick" in sys.argv else \
        [(sh, cg, bn) for sh in [(8480, 3072, 768), (8480, 768, 3072), (8480, 768, 768)] for cg in (1, 2) for bn in (128, 256)]
    for shape, cg, bn in cases:
        if True:
            if True:
                for dbg, what in [(0, "full"), (1, "no stores"), (3, "no stores, no tmem ld"), (4, "no MMA"), (12, "no MMA no TMA"), (8, "no TMA")]:
                    env = dict(os.environ, OFAB_GEMM_DBG=str(dbg), OFAB_GEMM_CG=str(cg), OFAB_GEMM_BN=str(bn))
                    r = subprocess.run([sys.executable, __file__, "child"] + [str(x) for x in shape], env=env, capture_output=True, text=True)
                    print(f"{shape} cg={cg} bn={bn} {what:24s} {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
