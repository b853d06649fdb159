"""GPU: CTC criterion kernels (csrc/ctc.cu) against the oracle (oracle/oracle_ctc.py, pinned to torch's
F.ctc_loss by tests/golden/ctc.pt) and against that fixture: fp32 logits -> nll 1e-5 relative, gradient 1e-4 rel-L2;
bf16 logits (what the CTC head GEMM emits) -> gradient stored in bf16, 6e-3."""
import os

import pytest
import torch

from oracle import oracle_ctc as oc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("time_major", [True, False])
def test_ctc_matches_torch_ctc_loss_fixture(time_major):
    from ofasys_b200 import ops

    fx = torch.load(os.path.join(GOLD, "ctc.pt"), weights_only=False)
    logits, targets, in_len, tgt_len, blank = oc.make_case()  # [T, B, C]
    x = (logits if time_major else logits.transpose(0, 1).contiguous()).cuda().requires_grad_(True)
    loss, nll = ops.ctc_loss_sum(x, targets.cuda(), tgt_len.cuda(), in_len.cuda(), blank=blank, zero_infinity=True, time_major=time_major)
    (2.0 * loss).backward()
    assert torch.isinf(nll[3]).item() and torch.isinf(fx["nll"][3]).item()
    assert torch.allclose(nll[:3].cpu().double(), fx["nll"][:3], rtol=1e-5)
    assert abs(loss.item() - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    g = x.grad if time_major else x.grad.transpose(0, 1)
    assert _rel(g, 2.0 * fx["dlogits"]) <= 1e-4
    assert not g[:, 3].any()  # zero_infinity: the impossible utterance contributes nothing
    assert not g[int(in_len[1]):, 1].any()  # frames past an utterance's length
    xo = logits.double().requires_grad_(True)
    lo, _ = oc.ctc_loss_sum(xo, targets, in_len, tgt_len, blank, True)
    (2.0 * lo).backward()
    assert _rel(g, xo.grad) <= 1e-4


def test_ctc_head_shape_bf16_asr_size():
    """ASR size: 248 encoder frames, B = 16, phone vocabulary 120, transcripts up to 60 labels, bf16
    logits as the CTC head GEMM produces them, batch-major, ragged frame counts."""
    from ofasys_b200 import ops

    g = torch.Generator().manual_seed(1)
    B, T, C, Lmax = 16, 248, 120, 60
    logits = (torch.randn(B, T, C, generator=g) * 2).to(torch.bfloat16)
    targets = torch.randint(2, C, (B, Lmax), generator=g)
    tgt_len = torch.randint(20, Lmax + 1, (B,), generator=g)
    in_len = torch.randint(180, T + 1, (B,), generator=g)
    x = logits.cuda().requires_grad_(True)
    loss, nll = ops.ctc_loss_sum(x, targets.cuda(), tgt_len.cuda(), in_len.cuda(), blank=1)
    loss.backward()
    xo = logits.float().transpose(0, 1).contiguous().requires_grad_(True)
    lo, no = oc.ctc_loss_sum(xo, targets, in_len, tgt_len, 1, True)
    lo.backward()
    assert torch.allclose(nll.cpu(), no.float(), rtol=2e-5)
    assert abs(loss.item() - lo.item()) <= 2e-5 * lo.item()
    assert _rel(x.grad.transpose(0, 1), xo.grad) <= 6e-3
    # a gradient row sums to zero over the classes (softmax minus a distribution over labels)
    assert x.grad.float().sum(-1).abs().max().item() <= 2e-2
