"""GPU: incremental decoding with K|V caches (reference: model/transformer.py:417-522 with
incremental_state, module/multihead_attention.py:188-279,393-409).

Oracle: for a causal decoder the features of position t computed step by step against cached keys / values are the
same numbers as row t of the full (teacher-forced) forward -- that full forward is the oracle path already pinned to the
reference (tests/golden).  So every step's logits are compared with (a) row t of the CUDA path's own full forward and
(b) row t of the oracle's fp32 logits, at the bf16 tolerances of tests/test_model_gpu.py; the greedy tokens must agree
with the oracle's wherever its top-2 margin is not within that noise.  Beam reordering is checked as a permutation
property."""
import os

import pytest
import torch

import ofasys_b200 as ob
from oracle import cases
from oracle import oracle_model as om
from util import bf16_round_state_dict, build_product, load_golden, rel_l2, to_product_slots

pytestmark = pytest.mark.gpu


def _setup(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    m = build_product(name)
    m.load_state_dict(sd, strict=False)
    m = m.to(torch.bfloat16).to(dev).eval()
    slots, target = cases.make_inputs(name)
    return dev, sd, m, slots, target


def _step_slots(pslots, t):
    """source slots unchanged; the target slot holds the prefix [:, :t+1] (what the sequence generator feeds)."""
    out = []
    for s in pslots:
        out.append(s if s.is_src else ob.Slot(s.modality, False, s.value[:, : t + 1].contiguous(), attributes=s.attributes))
    return out


@pytest.mark.parametrize("name", ["text_A", "patch_B", "text_B"])
def test_incremental_steps_equal_full_forward(name):
    dev, sd, m, slots, target = _setup(name)
    # decoding never feeds padding inside the prefix: use an unpadded target slot
    for s in slots:
        if not s.is_src:
            s.value = torch.where(s.value == om.PAD, torch.full_like(s.value, 5), s.value)
    pslots = to_product_slots(slots, dev)
    with torch.no_grad():
        full, _ = m(pslots)
        logits_ref, _ = om.model_forward(bf16_round_state_dict(sd), cases.oracle_cfg(name), slots)
        enc = m.encoder([s for s in pslots if s.is_src])
        T = full.shape[1]
        state = {}
        steps = []
        for t in range(T):
            lg, _ = m.decoder(_step_slots([s for s in pslots if not s.is_src], t), encoder_out=enc, incremental_state=state)
            assert tuple(lg.shape) == (full.shape[0], 1, full.shape[2])
            steps.append(lg[:, 0])
    inc = torch.stack(steps, dim=1)
    assert rel_l2(inc.float(), full.float()) <= 4e-3, rel_l2(inc.float(), full.float())
    assert rel_l2(inc.float(), logits_ref) <= 6e-3, rel_l2(inc.float(), logits_ref)
    gold = os.path.join(os.path.dirname(__file__), "golden", f"incr_{name}.pt")
    if os.path.exists(gold):  # the reference's OWN step-by-step logits (oracle/make_golden_incremental.py); == logits_ref to 1e-5
        ref_inc = torch.load(gold, weights_only=False)["incremental_logits"]
        assert rel_l2(inc.float(), ref_inc) <= 7e-3, rel_l2(inc.float(), ref_inc)
    top2 = logits_ref.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 0.05 * logits_ref.abs().amax(dim=-1)
    assert clear.float().mean() > 0.5
    assert torch.equal(inc.argmax(-1).cpu()[clear], logits_ref.argmax(-1)[clear])  # greedy choice = the oracle's
    # cache bookkeeping: one K|V row per step in every decoder self-attention, the encoder projection cached once
    for layer in m.decoder.layers:
        assert state[layer.self_attn._state_key]["len"] == T
        assert state[layer.encoder_attn._state_key]["kv"].shape[1] == enc["_encoder_out_bt"].shape[1]


def test_reorder_incremental_state_is_a_batch_permutation():
    dev, sd, m, slots, target = _setup("text_A")
    for s in slots:
        if not s.is_src:
            s.value = torch.where(s.value == om.PAD, torch.full_like(s.value, 5), s.value)
    pslots = to_product_slots(slots, dev)
    tgt = [s for s in pslots if not s.is_src]
    perm = torch.tensor([2, 0, 1], device=dev)
    with torch.no_grad():
        enc = m.encoder([s for s in pslots if s.is_src])
        state = {}
        for t in range(4):
            m.decoder(_step_slots(tgt, t), encoder_out=enc, incremental_state=state)
        ref5, _ = m.decoder(_step_slots(tgt, 4), encoder_out=enc, incremental_state=state)  # step 5 in the original order
        # same prefix decoded again, then the beams are permuted before step 5
        state2 = {}
        for t in range(4):
            m.decoder(_step_slots(tgt, t), encoder_out=enc, incremental_state=state2)
        m.decoder.reorder_incremental_state(state2, perm)
        enc2 = m.encoder.reorder_encoder_out(enc, perm)
        tgt2 = [ob.Slot(s.modality, False, s.value.index_select(0, perm), attributes=s.attributes) for s in tgt]
        got5, _ = m.decoder(_step_slots(tgt2, 4), encoder_out=enc2, incremental_state=state2)
    assert torch.equal(got5, ref5.index_select(0, perm))
