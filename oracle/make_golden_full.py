"""TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference) -- full-size goldens of BASELINE.json
configs[3] and configs[4] from the UNMODIFIED reference (oracle/ref_shim.py):

  cfg4_cotrain_base     OFA-base 12L/12L d=768, Mode A, ResNet-101: image_caption (224^2 + 8 tok -> 64 tok) + VQA (224^2 + 16
                        tok -> 8 tok) + text_infilling (128 -> 128), B=2 each, gradients ACCUMULATED over the three task batches
                        (engine/trainer.py:747-830 loops over the tasks of a step before the exchange)
  cfg5_large_video      OFA-large 24L/12L d=1024 H=16, ResNet-152, video_caption: 16 x 224^2 frames (S = 3136 + 8) -> 64 tok, B=1
  cfg5_large_grounding  the same model, visual_grounding: 512^2 image (S = 1024 + 16) -> BOX target (bos + 4 `<bin>` tokens), B=2

    python -m oracle.make_golden_full [case ...]

Each fixture holds reference OUTPUTS only: per task loss, logits at 64 fixed columns + logsumexp per position; for every
parameter the gradient's (sum, L1, L2) in fp64, the full tensor when it is small and bug-sensitive, `GRAD_SAMPLES` elements at
seeded positions and `GRAD_PROJ` seeded random projections (so a GPU test compares FULL gradient tensors -- positions and
signs -- without shipping 600 M numbers).  Weights / inputs are regenerated from seeds (oracle/cases.py).
"""
import os
import sys
import time
import zlib

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ref_shim  # noqa: E402
from oracle.make_golden import OUT, SMALL_FULL, digest, to_ref_slots  # noqa: E402

GRAD_SAMPLES = 1024
GRAD_PROJ = 2


def grad_probe_indices(name, numel):
    """Seeded element positions of parameter `name` (shared with the tests)."""
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randint(0, numel, (min(GRAD_SAMPLES, numel),), generator=g)


def grad_probe_vectors(name, numel):
    """Seeded +-1 projection vectors of parameter `name` (shared with the tests): [GRAD_PROJ, numel] int8."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ 0x5BD1E995) & 0x7FFFFFFF)
    return (torch.randint(0, 2, (GRAD_PROJ, numel), generator=g, dtype=torch.int8) * 2 - 1)


def grad_record(named_grads):
    stats, full, samples, proj = {}, {}, {}, {}
    for k, gr in named_grads.items():
        if gr is None:
            stats[k] = None
            continue
        gd = gr.detach().double().reshape(-1)
        stats[k] = torch.tensor([gd.sum().item(), gd.abs().sum().item(), gd.norm().item()], dtype=torch.float64)
        if any(t in k for t in SMALL_FULL) and gd.numel() <= 8192:
            full[k] = gr.detach().float().clone()
        samples[k] = gd[grad_probe_indices(k, gd.numel())].float()
        r = grad_probe_vectors(k, gd.numel())
        proj[k] = torch.stack([(gd * r[i].double()).sum() for i in range(GRAD_PROJ)])
    return stats, full, samples, proj


def run(name):
    t0 = time.time()
    c = cases.CASES[name]
    cfg = c["cfg"]
    torch.set_num_threads(8)
    m, ns = ref_shim.build_reference_model(
        arch="tiny", enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"], vocab=cfg["vocab"], adaptors=c["adaptors"], mode=cfg["mode"],
        dims=(cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"]), resnet_type=cfg.get("resnet_type"))
    assert m.cfg.adaptor.image_resnet.resnet_type == cfg["resnet_type"]
    spec = cases.param_spec_from_state_dict(m.state_dict())
    sd = cases.synth_state_dict(spec, seed=0)
    missing = torch.nn.Module.load_state_dict(m, sd, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("rp_bucket") or k.endswith("version") for k in missing.missing_keys)
    del sd
    m.train()
    out = {"case": name, "spec": spec, "tasks": []}
    for ti, (slots, target) in enumerate(cases.make_task_inputs(name)):
        logits, extra = m(to_ref_slots(ns, slots))
        lprobs = m.get_normalized_probs((logits, extra), log_probs=True).view(-1, logits.size(-1))
        loss = F.nll_loss(lprobs, target.view(-1), ignore_index=1, reduction="sum")
        loss.backward()  # accumulates over the tasks of the step
        lg = logits.detach().float()
        cols = torch.arange(0, lg.shape[-1], lg.shape[-1] // 64)[:64]
        out["tasks"].append({"loss": loss.detach().clone(), "ntokens": int((target != 1).sum()), "logit_cols": cols,
                             "logits_sampled": lg[..., cols].clone(), "lse": torch.logsumexp(lg, dim=-1)})
        print(f"  {name} task {ti}: loss {loss.item():.5f} logits {tuple(lg.shape)}  ({time.time() - t0:.0f} s)", flush=True)
        del logits, extra, lprobs, loss, lg
    out["grad_stats"], out["grad_full"], out["grad_samples"], out["grad_proj"] = grad_record({k: p.grad for k, p in m.named_parameters()})
    ints = {}
    for k, b in m.named_buffers():
        if k.endswith("rp_bucket"):
            ints[k] = (tuple(b.shape), digest(b))
    out["ints"] = ints
    path = os.path.join(OUT, f"{name}.pt")
    torch.save(out, path)
    print(f"{name}: {len(spec)} tensors, {sum(int(torch.tensor(s).prod()) for s in spec.values()) / 1e6:.1f} M values, file {os.path.getsize(path) / 1024:.0f} KiB, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    for n in sys.argv[1:] or ["cfg4_cotrain_base", "cfg5_large_grounding", "cfg5_large_video"]:
        run(n)
