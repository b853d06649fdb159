"""TEST INFRASTRUCTURE ONLY (container only) -- golden for the BOX-modality target (BASELINE configs[4]
visual_grounding): the unmodified reference on IMAGE (ResNet) + TEXT -> BOX, where the target slot holds `<bin>` tokens
from the integer quantisation of preprocessor/default/box.py:101-110 and is routed to the text adaptor
(adaptor/general.py:36-46).  Geometry and weights of case `resnet_A`.  Output: tests/golden/box_A.pt.
    python -m oracle.make_golden_box
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ref_shim  # noqa: E402
from oracle import oracle_model as om  # noqa: E402
from oracle.make_golden import OUT, to_ref_slots  # noqa: E402


def make_box_inputs(seed=99, B=4, V=512, num_bins=100):
    """Shared with the tests: 4 box coordinates -> bins -> `<bin>` ids in the last num_bins vocabulary entries."""
    g = torch.Generator().manual_seed(seed)
    coords = torch.rand(B, 4, generator=g) * 512
    bins = om.quantize_box(coords, 512, num_bins)
    first_bin = V - num_bins
    prev = torch.cat([torch.zeros(B, 1, dtype=torch.long), first_bin + bins], dim=1)  # bos + 4 bin tokens
    target = torch.cat([first_bin + bins, torch.full((B, 1), 2, dtype=torch.long)], dim=1)
    slots = [om.OSlot(om.IMAGE, True, torch.randn(B, 3, 64, 64, generator=g), adaptor="image_resnet"),
             om.OSlot(om.TEXT, True, cases._tokens(g, B, 16, V)), om.OSlot(om.BOX, False, prev)]
    return slots, target, coords, bins


def main():
    name = "resnet_A"
    c = cases.CASES[name]
    cfg = c["cfg"]
    m, ns = ref_shim.build_reference_model(
        arch="tiny", enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"], vocab=cfg["vocab"],
        adaptors=c["adaptors"], mode=cfg["mode"], dims=(cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"]),
    )
    spec = cases.param_spec_from_state_dict(m.state_dict())
    torch.nn.Module.load_state_dict(m, cases.synth_state_dict(spec, seed=0), strict=False)
    m.train()
    slots, target, coords, bins = make_box_inputs()
    logits, extra = m(to_ref_slots(ns, slots))
    lprobs = m.get_normalized_probs((logits, extra), log_probs=True).view(-1, logits.size(-1))
    loss = F.nll_loss(lprobs, target.view(-1), ignore_index=1, reduction="sum")
    path = os.path.join(OUT, "box_A.pt")
    torch.save({"logits": logits.detach().float().clone(), "loss": loss.detach().clone(), "bins": bins}, path)
    print(f"box_A: logits {tuple(logits.shape)} loss {loss.item():.5f}; wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
