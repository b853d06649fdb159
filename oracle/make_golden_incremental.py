"""TEST INFRASTRUCTURE ONLY (container only) -- golden vectors for incremental decoding from the
UNMODIFIED reference: its decoder is driven step by step with `incremental_state` exactly as its sequence generator does
(generator/sequence_generator.py:655-790 -> model/transformer.py:301-363,417-522; K/V caches of
module/multihead_attention.py:188-279) and the logits of every step are stored, next to the reference's own
teacher-forced full forward.  Weights: the seeded synthetic state dict rounded to bf16 (what the CUDA product holds).
Output: tests/golden/incr_<case>.pt.      python -m oracle.make_golden_incremental
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ref_shim  # noqa: E402
from oracle.make_golden import OUT, to_ref_slots  # noqa: E402


def run_case(name):
    c = cases.CASES[name]
    cfg = c["cfg"]
    m, ns = ref_shim.build_reference_model(
        arch="tiny", enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"], vocab=cfg["vocab"],
        adaptors=c["adaptors"], mode=cfg["mode"], dims=(cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"]),
    )
    spec = cases.param_spec_from_state_dict(m.state_dict())
    sd = cases.synth_state_dict(spec, seed=0)
    seen = {}
    sd_r = {}
    for k, v in sd.items():  # bf16-rounded, tied tensors stay tied
        if v.data_ptr() not in seen:
            seen[v.data_ptr()] = v.to(torch.bfloat16).float()
        sd_r[k] = seen[v.data_ptr()]
    torch.nn.Module.load_state_dict(m, sd_r, strict=False)
    m.eval()
    slots, _ = cases.make_inputs(name)
    for s in slots:
        if not s.is_src:  # decoding never feeds padding inside the prefix
            s.value = torch.where(s.value == 1, torch.full_like(s.value, 5), s.value)
    rslots = to_ref_slots(ns, slots)
    src = [s for s in rslots if s.is_src]
    tgt = [s for s in rslots if not s.is_src][0]
    with torch.no_grad():
        full, _ = m(rslots)
        enc = m.encoder(src)
        state = {}
        steps = []
        T = tgt.value.shape[1]
        for t in range(T):
            step_slot = ns.Slot(tgt.modality, False, tgt.value[:, : t + 1], attributes=None)
            lg, _ = m.decoder([step_slot], encoder_out=enc, incremental_state=state)
            assert lg.shape[1] == 1, lg.shape
            steps.append(lg[:, 0].float())
    inc = torch.stack(steps, dim=1)
    err = ((inc - full.float()).norm() / full.float().norm()).item()
    path = os.path.join(OUT, f"incr_{name}.pt")
    torch.save({"case": name, "incremental_logits": inc.clone(), "full_logits": full.float().clone()}, path)
    print(f"{name}: reference incremental vs its own full forward rel-L2 {err:.2e}; wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    for n in sys.argv[1:] or ["text_A", "patch_B"]:
        run_case(n)
