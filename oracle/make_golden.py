"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
For every case in oracle/cases.py: build the reference GeneralistModel (via oracle/ref_shim.py),
overwrite its parameters with the seeded synthetic weights, run the reference forward + the
reference criterion's loss (sum-CE over non-pad targets) + backward, and store
  spec      : {param name: shape}  (lets tests regenerate the same weights without the reference)
  logits    : full [B,T,V] fp32 when V is small, else `lse` [B,T] + logits at 64 fixed columns
  loss      : scalar
  grads     : per parameter (sum, L1, L2) + full tensors for the small, bug-sensitive ones
  ints      : integer artefacts that must be bit-exact (bucket tables digests, pad masks)
The fixtures hold reference OUTPUTS only (no reference source).
"""
import hashlib
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SMALL_FULL = ("c_attn", "rel_pos_table", "layer_norm", "layernorm", "_ln.", "type_embedding", "cls_token", ".bias")


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def to_ref_slots(ns, slots):
    out = []
    for s in slots:
        mod = ns.ModalityType(s.modality)
        attrs = f"adaptor={s.adaptor}" if s.adaptor else None
        out.append(ns.Slot(mod, s.is_src, s.value, attributes=attrs))
    return out


def run_case(name):
    c = cases.CASES[name]
    cfg = c["cfg"]
    torch.set_num_threads(8)
    m, ns = ref_shim.build_reference_model(
        arch="tiny", enc_layers=cfg["enc_layers"], dec_layers=cfg["dec_layers"], vocab=cfg["vocab"],
        adaptors=c["adaptors"], mode=cfg["mode"], dims=(cfg["embed_dim"], cfg["heads"], cfg["ffn_dim"]),
    )
    if "resnet_type" in cfg:
        assert m.cfg.adaptor.image_resnet.resnet_type == cfg["resnet_type"]
    ref_sd = m.state_dict()
    spec = cases.param_spec_from_state_dict(ref_sd)
    sd = cases.synth_state_dict(spec, seed=0)
    missing = torch.nn.Module.load_state_dict(m, sd, strict=False)  # skip the ckpt-upgrade wrapper (fairseq_model.py:105-127)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all(k.endswith("rp_bucket") or k.endswith("version") for k in missing.missing_keys), missing.missing_keys
    m.train()
    slots, target = cases.make_inputs(name)
    logits, extra = m(to_ref_slots(ns, slots))
    # reference criterion (engine/criterion/cross_entropy.py:62-67): fp32 log-softmax + sum NLL, ignore pad
    lprobs = m.get_normalized_probs((logits, extra), log_probs=True).view(-1, logits.size(-1))
    loss = F.nll_loss(lprobs, target.view(-1), ignore_index=1, reduction="sum")
    loss.backward()
    g = {"case": name, "spec": spec, "loss": loss.detach().clone(), "ntokens": int((target != 1).sum())}
    lg = logits.detach().float()
    if lg.numel() <= 1 << 18:
        g["logits"] = lg.clone()
    else:
        cols = torch.arange(0, lg.shape[-1], lg.shape[-1] // 64)[:64]
        g["logit_cols"] = cols
        g["logits_sampled"] = lg[..., cols].clone()
        g["lse"] = torch.logsumexp(lg, dim=-1)
    grads = {}
    full = {}
    seen = set()
    for k, p in m.named_parameters():
        if p.grad is None:
            grads[k] = None
            continue
        gr = p.grad.detach().double()  # stats in fp64: fp32 norms of 1e7-element grads lose 1e-3
        grads[k] = torch.tensor([gr.sum().item(), gr.abs().sum().item(), gr.norm().item()], dtype=torch.float64)
        if any(t in k for t in SMALL_FULL) and gr.numel() <= 8192:
            full[k] = gr.float().clone()
    g["grad_stats"] = grads
    g["grad_full"] = full
    if lg.numel() > 1 << 18:  # full-size configs: probes of the FULL gradient tensors (seeded samples + random projections)
        g["grad_samples"], g["grad_proj"] = {}, {}
        for k, p in m.named_parameters():
            if p.grad is not None:
                g["grad_samples"][k], g["grad_proj"][k] = cases.grad_probes(k, p.grad)
    ints = {}
    for k, b in m.named_buffers():
        if k.endswith("rp_bucket"):
            ints[k] = (tuple(b.shape), digest(b), b[:64, :64].clone() if b.dim() == 2 else None)
    g["ints"] = ints
    torch.save(g, os.path.join(OUT, f"{name}.pt"))
    sz = os.path.getsize(os.path.join(OUT, f"{name}.pt"))
    print(f"{name}: loss={loss.item():.6f} ntok={g['ntokens']} params={len(spec)} file={sz/1024:.0f} KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(cases.CASES)
    for n in names:
        run_case(n)
