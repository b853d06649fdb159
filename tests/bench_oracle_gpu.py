"""DIAGNOSTIC (not a bench line): the oracle port of the reference algorithm run EAGERLY ON THE GPU in bf16 -- torch's own
kernels (cuBLAS GEMMs, ATen softmax / layer_norm / elementwise), i.e. what the reference's stock PyTorch path costs on the same
B200 for the headline workload (image_caption OFA-base 12L/12L, 224^2 patch-embed + 8-tok prompt -> 64-tok caption, B=64).
Lives under tests/ because it executes oracle/ (test infrastructure); results are copied into profiles/.

    python tests/bench_oracle_gpu.py [batch]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definition: CFG, host_batch, _spec_cache)
from oracle import cases  # noqa: E402
from oracle import oracle_model as om  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
cfg = om.OracleConfig(**bench.CFG)
sd = cases.synth_state_dict(bench._spec_cache(), seed=0)
out = {}
for dtype, tag in ((torch.bfloat16, "bf16"), (torch.float32, "fp32_tf32")):
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    sdd = {}
    seen = {}
    for k, v in sd.items():
        if v.data_ptr() not in seen:
            seen[v.data_ptr()] = v.to(dev, dtype) if v.is_floating_point() else v.to(dev)
        sdd[k] = seen[v.data_ptr()]
    hb = bench.host_batch(B, 1234, pin=False)
    slots = [om.OSlot(om.IMAGE, True, hb["img"].to(dev, dtype), adaptor="image_patch_embed"), om.OSlot(om.TEXT, True, hb["prompt"].to(dev)),
             om.OSlot(om.TEXT, False, hb["prev"].to(dev))]
    tgt = hb["tgt"].to(dev)
    with torch.device(dev):
        for _ in range(3):
            om.loss_and_grads(sdd, cfg, slots, tgt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            om.loss_and_grads(sdd, cfg, slots, tgt)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out[tag] = {"ms_per_step": ms, "seq_per_s": B / (ms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    del sdd, seen
    torch.cuda.empty_cache()
print(json.dumps({"diagnostic": "oracle port of the reference algorithm, torch eager on the GPU (cuBLAS + ATen kernels)", "workload": bench.WORKLOAD, "per_gpu_batch": B, **out}))
