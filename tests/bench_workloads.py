"""fwd+bwd throughput of the other BASELINE.json workloads on one GPU (bench.py measures configs[1] only; these are extra
measured lines for profiles/, same method: inputs resident, whole step replayed as one CUDA graph, CUDA events).

  python tests/bench_workloads.py asr [B]              # configs[2]: OFA-base ASR, fbank [B,998,80] + 12-tok prompt -> 128 tok (Mode A)
  python tests/bench_workloads.py caption_resnet [B]   # configs[1] variant 2b: ResNet-101 224^2 + 8-tok prompt -> 64 tok (Mode A)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/: it uses the oracle's case definitions)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ofasys_b200 as ob  # noqa: E402
from oracle import cases  # noqa: E402
from util import build_product  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "asr"
V = 50265
base = dict(embed_dim=768, heads=12, ffn_dim=3072, enc_layers=12, dec_layers=12, vocab=V, mode="A")
if which == "asr":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cases.CASES["bench"] = dict(cfg=base, adaptors=("text", "audio"), kind="audio", B=B, S=12, T=128, L=998)
    gflop = 445.52  # 3 x F_fwd per sequence, SURVEY 8d
elif which == "caption_resnet":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cases.CASES["bench"] = dict(cfg=dict(base, resnet_type="resnet101"), adaptors=("text", "image_resnet"), kind="resnet", B=B, S=8, T=64, image=224)
    gflop = 226.64
else:
    raise SystemExit(which)
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = build_product("bench").to(torch.bfloat16).to(dev).train()
g = torch.Generator().manual_seed(1234)
MT = ob.ModalityType
prev = torch.randint(4, V, (B, cases.CASES["bench"]["T"]), generator=g)
prev[:, 0] = 0
prev[: max(1, B // 10), -cases.CASES["bench"]["T"] // 4:] = 1
tgt = torch.roll(prev, -1, 1)
tgt[:, -1] = 2
tgt[prev == 1] = 1
tgt[torch.roll(prev == 1, -1, 1)] = 1
prompt = torch.randint(4, V, (B, cases.CASES["bench"]["S"]), generator=g)
if which == "asr":
    fb = torch.randn(B, 998, 80, generator=g)
    lens = torch.randint(700, 999, (B,), generator=g)
    src = ob.Slot(MT.AUDIO, True, {"fbank": fb.to(dev), "fbank_lengths": lens.to(dev)})
else:
    src = ob.Slot(MT.IMAGE, True, torch.randn(B, 3, 224, 224, generator=g).to(dev), attributes="adaptor=image_resnet")
slots = [src, ob.Slot(MT.TEXT, True, prompt.to(dev)), ob.Slot(MT.TEXT, False, prev.to(dev))]
tgt = tgt.to(dev)
params = [p for p in m.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    loss = m.forward_loss(slots, tgt)
    loss.backward()
    return loss


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
from ofasys_b200 import _lib  # noqa: E402

graph = True
try:
    c0 = _lib.launch_count
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        loss = step()
    launches = _lib.launch_count - c0
    run = gr.replay
except Exception as ex:
    print(f"[bench_workloads] graph capture failed ({type(ex).__name__}: {ex}); eager", file=sys.stderr)
    torch.cuda.synchronize()
    graph, launches, run = False, None, step
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
tf = B / (ms * 1e-3) * gflop / 1e3
print(json.dumps({"workload": which, "per_gpu_batch": B, "ms_per_step": ms, "seq_per_s": B / (ms * 1e-3), "cuda_graph": graph, "gpu_launches": launches,
                  "algorithmic_gflop_per_seq": gflop, "model_tflops": tf, "model_frac_of_bf16_peak": tf / pk["bf16_tflops_sustained"],
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
