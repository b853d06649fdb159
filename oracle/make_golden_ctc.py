"""TEST INFRASTRUCTURE ONLY -- golden vectors for the CTC term: torch's own F.ctc_loss (the third-party function the
reference calls, speech_to_text_loss.py:364-373) on oracle_ctc.make_case(), in float64.  Output: tests/golden/ctc.pt.
    python -m oracle.make_golden_ctc
"""
import os

import torch
import torch.nn.functional as F

from . import oracle_ctc as oc


def main():
    logits, targets, in_len, tgt_len, blank = oc.make_case()
    x = logits.double().requires_grad_(True)
    lprobs = torch.log_softmax(x, dim=-1)
    flat = torch.cat([targets[b, : int(tgt_len[b])] for b in range(targets.shape[0])])  # masked_select form (:355-357)
    per = F.ctc_loss(lprobs, flat, in_len, tgt_len, blank=blank, reduction="none", zero_infinity=False)
    loss = F.ctc_loss(lprobs, flat, in_len, tgt_len, blank=blank, reduction="sum", zero_infinity=True)
    loss.backward()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ctc.pt")
    torch.save({"nll": per.detach(), "loss": loss.detach(), "dlogits": x.grad.clone()}, path)
    print("wrote", path, os.path.getsize(path), "bytes; nll", per.tolist(), "loss", float(loss))


if __name__ == "__main__":
    main()
