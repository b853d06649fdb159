"""TEST INFRASTRUCTURE ONLY (container only: needs /root/reference) -- golden vectors for the label-smoothed criterion.

Runs the reference's own `label_smoothed_nll_loss` (engine/criterion/label_smoothed_cross_entropy.py:62-92), imported
UNMODIFIED through stub packages, on the seeded case of oracle_model.make_ls_case(); the fp32 log-softmax and the
padding filter around it are written out as in `compute_loss` (:175-191) / `get_lprobs_and_target` (:141-173).
Output: tests/golden/ls_ce.pt.      python -m oracle.make_golden_criterion
"""
import importlib
import os
import sys
import types

import torch

from . import oracle_model as om
from . import ref_shim


def _reference():
    ref_shim.install()
    for name, path in (("ofasys.engine", []), ("ofasys.engine.criterion", [os.path.join(ref_shim.REF_PKG, "engine", "criterion")])):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = path
            sys.modules[name] = m
    lg = types.ModuleType("ofasys.logging")
    lg.metrics = types.ModuleType("ofasys.logging.metrics")
    sys.modules["ofasys.logging"], sys.modules["ofasys.logging.metrics"] = lg, lg.metrics
    mod = importlib.import_module("ofasys.engine.criterion.label_smoothed_cross_entropy")
    return mod.label_smoothed_nll_loss


def main():
    fn = _reference()
    logits, target, eps = om.make_ls_case()
    x = logits.float().requires_grad_(True)
    lprobs = torch.log_softmax(x, dim=-1)  # get_normalized_probs(log_probs=True), ofa.py:287-299 (fp32)
    keep = target != om.PAD  # :178-180
    loss, nll, ntokens = fn(lprobs[keep], target[keep], eps, update_num=0)
    loss.backward()
    out = {"loss": loss.detach(), "nll_loss": nll.detach(), "ntokens": ntokens, "dlogits": x.grad.clone()}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ls_ce.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), "nll", float(nll), "ntokens", ntokens)
    # ---- constraint masks / constraint_range / drop-worst: the reference function on the reference's own masking recipe
    # (get_constraint_masks :147-157, get_lprobs_and_target :159-173, compute_loss :175-191 written out)
    import math

    logits, target, masks, (lo, hi) = om.make_constraint_case()
    res = {}
    for tag, use_masks, use_range, dw in (("range", False, True, 0.0), ("masks", True, False, 0.0), ("both", True, True, 0.0), ("dropworst", False, True, 0.25)):
        x = logits.float().requires_grad_(True)
        cm = None
        if use_range:
            cm = torch.ones(x.shape, dtype=torch.bool)
            cm[..., 4:lo] = 0
            cm[..., hi:] = 0
            if use_masks:
                cm = torch.logical_and(masks, cm)
        elif use_masks:
            cm = masks
        xm = x.masked_fill(~cm, -math.inf)
        lprobs = torch.log_softmax(xm, dim=-1).view(-1, x.size(-1))
        tgt = target.view(-1)
        cmf = cm.reshape(-1, cm.size(-1))
        keep = tgt != om.PAD
        l, n, nt = fn(lprobs[keep], tgt[keep], 0.1, update_num=5, drop_worst_ratio=dw, drop_worst_after=2, constraint_masks=cmf[keep])
        l.backward()
        res[tag] = {"loss": l.detach(), "nll_loss": n.detach(), "ntokens": nt, "dlogits": x.grad.clone()}
        print(tag, float(l), float(n), nt)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ls_ce_constraints.pt")
    torch.save(res, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
