"""TEST INFRASTRUCTURE ONLY (container only) -- pins the PLACES where the oracle applies dropout / drop-path (row A17,
oracle_model.DROP_HOOK) to the unmodified reference: the reference model runs in training mode with dropout 0.1,
attention dropout 0.05, activation dropout 0.15 and drop-path 0.2 while `torch.nn.functional.dropout` and `torch.rand`
are replaced by draws from ONE seeded generator (MaskStream); the oracle, fed the same stream through DROP_HOOK, must
reproduce the reference's logits and loss -- which it can only do if it draws masks of the same shapes in the same
order, i.e. applies them at the same call sites.  Output: tests/golden/drop_<case>.pt.
    python -m oracle.make_golden_dropout
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ref_shim  # noqa: E402
from oracle.make_golden import OUT, to_ref_slots  # noqa: E402

P_RES, P_ATTN, P_ACT, P_PATH = 0.1, 0.05, 0.15, 0.2
_TORCH_RAND = torch.rand  # the real one: MaskStream keeps drawing from it while torch.rand is patched


class MaskStream:
    """The shared source of randomness: every dropout / drop-path call draws its uniform numbers from here, in call order."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.calls = 0

    def uniform(self, shape):
        self.calls += 1
        return _TORCH_RAND(tuple(shape), generator=self.g)

    def dropout_mult(self, shape, p):  # F.dropout: zero with probability p, scale the rest by 1 / (1 - p)
        return (self.uniform(shape) >= p).float() / (1.0 - p)


def run_case(name, seed=77):
    c = cases.CASES[name]
    cfg = c["cfg"]
    m, ns = ref_shim.build_reference_model(
        arch="tiny", enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"], vocab=cfg["vocab"],
        adaptors=c["adaptors"], mode=cfg["mode"], dims=(cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"]),
    )
    spec = cases.param_spec_from_state_dict(m.state_dict())
    torch.nn.Module.load_state_dict(m, cases.synth_state_dict(spec, seed=0), strict=False)
    m.train()
    n_drop = n_path = 0
    for mod_name, mod in m.named_modules():
        cls = type(mod).__name__
        if cls == "Dropout":
            if mod_name.endswith("self_attn.dropout_module") or mod_name.endswith("encoder_attn.dropout_module"):
                mod.p = P_ATTN
                if cfg["mode"] == "B" and mod_name.startswith("encoder."):
                    # Mode B encoder self-attention takes F.multi_head_attention_forward (multihead_attention.py:155-186), whose
                    # dropout runs inside torch's fused attention with its own RNG: not replayable -> off for this fixture
                    mod.p = 0.0
            elif mod_name.endswith("activation_dropout_module"):
                mod.p = P_ACT
            else:
                mod.p = P_RES
            n_drop += 1
        elif cls == "DropPath":
            mod.drop_prob = P_PATH
            n_path += 1
    stream = MaskStream(seed)
    orig_dropout, orig_rand = F.dropout, torch.rand

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        return x * stream.dropout_mult(x.shape, p).to(x.dtype)

    def fake_rand(*shape, **kw):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (list, tuple, torch.Size)) else shape
        return stream.uniform(shape).to(kw.get("dtype") or torch.float32)

    slots, target = cases.make_inputs(name)
    F.dropout, torch.rand = fake_dropout, fake_rand
    try:
        logits, extra = m(to_ref_slots(ns, slots))
    finally:
        F.dropout, torch.rand = orig_dropout, orig_rand
    lprobs = m.get_normalized_probs((logits, extra), log_probs=True).view(-1, logits.size(-1))
    loss = F.nll_loss(lprobs, target.view(-1), ignore_index=1, reduction="sum")
    path = os.path.join(OUT, f"drop_{name}.pt")
    torch.save({"case": name, "seed": seed, "p": (P_RES, P_ATTN, P_ACT, P_PATH), "logits": logits.detach().float().clone(),
                "loss": loss.detach().clone(), "draws": stream.calls}, path)
    print(f"{name}: {n_drop} Dropout / {n_path} DropPath modules, {stream.calls} draws, loss {loss.item():.5f}; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    for n in sys.argv[1:] or ["text_A", "patch_B"]:
        run_case(n)
