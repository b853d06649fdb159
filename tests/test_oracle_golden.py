"""CPU: the oracle restatement (oracle/oracle_model.py) reproduces the outputs of the reference
code itself (fixtures made by oracle/make_golden.py from the unmodified reference files)."""
import hashlib
import os

import pytest
import torch

from oracle import cases
from oracle import oracle_model as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = ["text_A", "text_B", "patch_B", "audio_A", "resnet_A", "video_A", "large_A"]


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def load(name):
    return torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_outputs(name):
    g = load(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    cfg = cases.oracle_cfg(name)
    slots, target = cases.make_inputs(name)
    loss, logits, grads = om.loss_and_grads(sd, cfg, slots, target)
    assert int((target != 1).sum()) == g["ntokens"]
    # tolerances: fp32 CPU vs fp32 CPU, different op order only
    assert rel_l2(logits, g["logits"]) <= 1e-5
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    for k, st in g["grad_stats"].items():
        if st is None:
            continue
        gr = grads[k].double()
        l2 = gr.norm().item()
        # k_proj.bias grads are mathematically 0 (softmax shift invariance): absolute floor 1e-6
        assert abs(l2 - st[2].item()) <= 2e-4 * st[2].item() + 1e-6, (k, l2, st[2].item())
    for k, full in g["grad_full"].items():
        if full.norm() <= 1e-5:
            assert grads[k].norm() <= 2e-5, k
        else:
            assert rel_l2(grads[k], full) <= 2e-4, k


@pytest.mark.parametrize("name", SMALL)
def test_bucket_tables_bit_exact(name):
    """Integer position machinery must be bit-exact."""
    g = load(name)
    cfg = cases.oracle_cfg(name)
    for k, (shape, dig, corner) in g["ints"].items():
        if "token_rp_bucket" in k:
            t = om.make_token_bucket_position(cfg.token_bucket_size, cfg.max_position)
        elif "audio_rp_bucket" in k:
            t = om.make_token_bucket_position(cfg.max_position, 4096)
        elif "image_rp_bucket" in k:
            t = om.make_image_bucket_position(cfg.image_bucket_size, (2 * cfg.image_bucket_size - 1) ** 2 + 3)
        else:
            continue
        assert tuple(t.shape) == shape
        assert t.dtype == torch.int64
        assert hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest() == dig, k


def test_audio_length_quirk():
    # subsample.py:37-41: L=700 -> 175 (conv really yields 174); L=998 -> 250 reported vs 248 frames
    assert om.audio_out_lengths(torch.tensor([700, 998, 200, 150])).tolist() == [175, 250, 50, 38]


def test_box_quantisation():
    # box.py:101-110
    x = torch.tensor([0.0, 255.6, 511.0, 512.0])
    assert om.quantize_box(x).tolist() == [0, 499, 997, 999]


@pytest.mark.parametrize("name", ["cfg1_tiny", "cfg2_base", "cfg3_asr_base"])
def test_full_size_configs(name):
    """BASELINE.json configs[0..2] at their full model sizes (checksummed: V=50265 logits do not fit a fixture):
    cfg1 OFA-tiny 4L/4L d=256 text_infilling S=T=128; cfg2 OFA-base 12L/12L image_caption (patch-embed, 257+8 -> 64);
    cfg3 OFA-base ASR (fbank 998x80 ragged + 12-token prompt -> 128), batch 2 each."""
    g = load(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    cfg = cases.oracle_cfg(name)
    slots, target = cases.make_inputs(name)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    loss, logits, grads = om.loss_and_grads(sd, cfg, slots, target)
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    assert rel_l2(torch.logsumexp(logits, -1), g["lse"]) <= 1e-5
    assert rel_l2(logits[..., g["logit_cols"]], g["logits_sampled"]) <= 1e-5
    for k, st in g["grad_stats"].items():
        if st is not None:
            assert abs(grads[k].double().norm().item() - st[2].item()) <= 2e-4 * st[2].item() + 1e-6, k


@pytest.mark.parametrize("name", ["cfg2_base", "cfg3_asr_base"])
def test_full_gradient_probes_of_the_oracle(name):
    """the FULL gradient tensors of the oracle against the reference's seeded samples / projections (positions and signs)"""
    g = load(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    slots, target = cases.make_inputs(name)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    _, _, grads = om.loss_and_grads(sd, cases.oracle_cfg(name), slots, target)
    _check_probes(g, grads)


def _check_probes(g, grads):
    n = 0
    tot = sum(st[2].item() ** 2 for st in g["grad_stats"].values() if st is not None) ** 0.5
    for k, ref_s in g["grad_samples"].items():
        norm = g["grad_stats"][k][2].item()
        if norm <= 1e-6 * tot:  # k_proj.bias gradients are mathematically 0 (softmax shift invariance): rounding noise
            continue
        smp, prj = cases.grad_probes(k, grads[k])
        rms = max(norm / grads[k].numel() ** 0.5, 1e-12)
        assert ((smp.double() - ref_s.double()).pow(2).mean().sqrt() / rms).item() <= 1e-3, k
        assert ((prj - g["grad_proj"][k]).abs().max() / max(norm, 1e-9)).item() <= 1e-3, k
        n += 1
    assert n > 50


@pytest.mark.parametrize("name", ["cfg4_cotrain_base"] + (["cfg5_large_grounding", "cfg5_large_video"] if os.environ.get("OFAB_SLOW_TESTS") == "1" else []))
def test_full_size_multitask_oracle(name):
    """BASELINE.json configs[3] (co-training: caption + VQA + text_infilling batches of one step, gradients accumulated) at
    full OFA-base size against the dump of the unmodified reference (oracle/make_golden_full.py).  The OFA-large configs[4]
    cases (24L/12L, S = 1040 / 3144; minutes and tens of GB on CPU) run with OFAB_SLOW_TESTS=1."""
    g = load(name)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    cfg = cases.oracle_cfg(name)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    total = None
    for (slots, target), gt in zip(cases.make_task_inputs(name), g["tasks"]):
        loss, logits, grads = om.loss_and_grads(sd, cfg, slots, target)
        assert int((target != 1).sum()) == gt["ntokens"]
        assert abs(loss.item() - gt["loss"].item()) <= 1e-5 * abs(gt["loss"].item())
        assert rel_l2(torch.logsumexp(logits, -1), gt["lse"]) <= 1e-5
        assert rel_l2(logits[..., gt["logit_cols"]], gt["logits_sampled"]) <= 2e-5
        total = grads if total is None else {k: total[k] + grads[k] for k in total}
    for k, st in g["grad_stats"].items():
        if st is not None:
            assert abs(total[k].double().norm().item() - st[2].item()) <= 2e-4 * st[2].item() + 1e-6, k
    _check_probes(g, total)


def test_optimizer_oracle_matches_reference_adam():
    """oracle_optim.update (restatement of multiply_grads -> clip_grad_norm -> fp32 Adam -> bf16 copy) against the
    outputs of the reference's own Adam.step / clip_grad_norm_ (tests/golden/optim_adam.pt, oracle/make_golden_optim.py)."""
    from oracle import oracle_optim as oo

    fx = torch.load(os.path.join(GOLD, "optim_adam.pt"), weights_only=False)
    params, steps, hyper, scales = oo.make_case()
    masters = [p.float() for p in params]
    ms = [torch.zeros_like(m) for m in masters]
    vs = [torch.zeros_like(m) for m in masters]
    for k, (gs, c) in enumerate(zip(steps, scales)):
        norm, p16 = oo.update(masters, gs, ms, vs, k + 1, hyper["lr"], hyper["betas"], hyper["eps"], hyper["weight_decay"], c, hyper["max_norm"])
        assert abs(float(norm) - float(fx["norms"][k])) <= 1e-6 * float(fx["norms"][k])
        for name, cur in (("masters", masters), ("exp_avg", ms), ("exp_avg_sq", vs), ("params_bf16", p16)):
            for i in fx["small"]:
                assert torch.equal(cur[i], fx[name][k][i]), (name, k, i)  # same torch ops in the same order: bit-exact
            for i, t in enumerate(cur):
                s, a = t.double().sum(), t.double().abs().sum()
                assert abs(s - fx[name + "_sum"][k][i]) <= 1e-9 * max(1.0, float(a)), (name, k, i)
                assert abs(a - fx[name + "_abs"][k][i]) <= 1e-9 * max(1.0, float(a)), (name, k, i)


def test_label_smoothed_criterion_oracle_matches_reference():
    """oracle_model.label_smoothed_cross_entropy_sum against the reference's own label_smoothed_nll_loss
    (tests/golden/ls_ce.pt, oracle/make_golden_criterion.py)."""
    fx = torch.load(os.path.join(GOLD, "ls_ce.pt"), weights_only=False)
    logits, target, eps = om.make_ls_case()
    x = logits.float().requires_grad_(True)
    loss, nll, ntok = om.label_smoothed_cross_entropy_sum(x, target, eps)
    loss.backward()
    assert ntok == fx["ntokens"]
    assert abs(float(loss) - float(fx["loss"])) <= 1e-6 * abs(float(fx["loss"]))
    assert abs(float(nll) - float(fx["nll_loss"])) <= 1e-6 * abs(float(fx["nll_loss"]))
    assert ((x.grad - fx["dlogits"]).norm() / fx["dlogits"].norm()).item() <= 1e-6
    # eps = 0 is the plain criterion
    l0, n0, _ = om.label_smoothed_cross_entropy_sum(logits.float(), target, 0.0)
    assert abs(float(l0) - float(om.cross_entropy_sum(logits.float(), target))) <= 1e-6 * abs(float(l0)) and float(l0) == float(n0)


def test_constrained_criterion_oracle_matches_reference():
    """oracle_model.constrained_criterion (constraint_range / constraint masks / drop-worst, label_smoothed_cross_entropy.py:62-92,
    147-191) against the reference's own label_smoothed_nll_loss on its own masking recipe (tests/golden/ls_ce_constraints.pt)."""
    fx = torch.load(os.path.join(GOLD, "ls_ce_constraints.pt"), weights_only=False)
    logits, target, masks, rng = om.make_constraint_case()
    for tag, use_masks, use_range, dw in (("range", False, True, 0.0), ("masks", True, False, 0.0), ("both", True, True, 0.0), ("dropworst", False, True, 0.25)):
        x = logits.float().requires_grad_(True)
        loss, nll, ntok = om.constrained_criterion(x, target, 0.1, rng if use_range else None, masks if use_masks else None, update_num=5,
                                                   drop_worst_ratio=dw, drop_worst_after=2)
        loss.backward()
        f = fx[tag]
        assert ntok == f["ntokens"], tag
        assert abs(float(loss) - float(f["loss"])) <= 1e-6 * abs(float(f["loss"])), tag
        assert abs(float(nll) - float(f["nll_loss"])) <= 1e-6 * abs(float(f["nll_loss"])), tag
        assert ((x.grad - f["dlogits"]).norm() / f["dlogits"].norm()).item() <= 1e-6, tag


def test_audio_front_end_oracle_matches_torchaudio_and_reference_cmvn():
    """oracle_audio.fbank against tests/golden/fbank.pt (torchaudio.compliance.kaldi.fbank, the function the reference
    calls, + the reference's own UtteranceCMVN) and, when torchaudio is importable, against torchaudio directly.
    Same torch ops in the same order -> 1e-6."""
    from oracle import oracle_audio as oa

    fx = torch.load(os.path.join(GOLD, "fbank.pt"), weights_only=False)
    wav, lengths = oa.make_case()
    for b in range(wav.shape[0]):
        f = oa.fbank(wav[b:b + 1, : int(lengths[b])])
        assert f.shape == fx["fbank"][b].shape
        assert (f - fx["fbank"][b]).abs().max().item() <= 1e-5
        c = torch.from_numpy(oa.utterance_cmvn(f.numpy()))
        assert (c - fx["cmvn"][b]).abs().max().item() <= 1e-4
    try:
        import torchaudio.compliance.kaldi as ta_kaldi
    except Exception:
        return
    g = torch.Generator().manual_seed(3)
    w = torch.randn(1, 16000, generator=g) * 2000
    assert (oa.fbank(w) - ta_kaldi.fbank(w, num_mel_bins=80, sample_frequency=16000)).abs().max().item() <= 1e-5
    # the product's host-side tables are the same numbers
    from ofasys_b200.preprocessor.audio import Fbank, kaldi_mel_banks, kaldi_window

    assert torch.equal(kaldi_window(400), torch.hann_window(400, periodic=False).pow(0.85))
    mel = torch.nn.functional.pad(oa.mel_banks(80, 512, 16000.0).to(torch.float32), (0, 1))
    assert torch.equal(kaldi_mel_banks(80, 512, 16000.0), mel)
    assert Fbank().num_frames(16000) == 98 and Fbank().num_frames(399) == 0


def test_ctc_oracle_matches_torch_ctc_loss():
    """oracle_ctc (restated alpha recursion, autograd gradient) against F.ctc_loss's own outputs (tests/golden/ctc.pt) and
    against F.ctc_loss directly, float64."""
    import torch.nn.functional as F

    from oracle import oracle_ctc as oc

    fx = torch.load(os.path.join(GOLD, "ctc.pt"), weights_only=False)
    logits, targets, in_len, tgt_len, blank = oc.make_case()
    x = logits.double().requires_grad_(True)
    loss, nll = oc.ctc_loss_sum(x.to(torch.float64), targets, in_len, tgt_len, blank, zero_infinity=True)
    loss.backward()
    per = oc.ctc_nll(torch.log_softmax(logits.double(), -1), targets, in_len, tgt_len, blank)
    assert torch.isinf(per[3]) and torch.isinf(fx["nll"][3])  # 4 frames cannot carry 8 labels
    assert torch.allclose(per[:3], fx["nll"][:3], rtol=1e-10, atol=1e-10)
    assert abs(float(loss) - float(fx["loss"])) <= 1e-9 * abs(float(fx["loss"]))
    assert ((x.grad - fx["dlogits"]).norm() / fx["dlogits"].norm()).item() <= 1e-9
    flat = torch.cat([targets[b, : int(tgt_len[b])] for b in range(4)])
    direct = F.ctc_loss(torch.log_softmax(logits.double(), -1), flat, in_len, tgt_len, blank=blank, reduction="sum", zero_infinity=True)
    assert abs(float(direct) - float(loss)) <= 1e-9 * abs(float(direct))


@pytest.mark.parametrize("name", ["text_A", "patch_B"])
def test_incremental_decoding_golden_pins_the_oracle(name):
    """tests/golden/incr_<case>.pt holds the logits the UNMODIFIED reference produced when driven step by step with
    incremental_state (oracle/make_golden_incremental.py).  They equal the reference's own teacher-forced forward (the
    property the GPU test of incremental decoding relies on) and the oracle's forward on the same inputs."""
    from util import bf16_round_state_dict

    fx = torch.load(os.path.join(GOLD, f"incr_{name}.pt"), weights_only=False)
    inc, full = fx["incremental_logits"], fx["full_logits"]
    assert ((inc - full).norm() / full.norm()).item() <= 1e-5
    g = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    sd = bf16_round_state_dict(cases.synth_state_dict(g["spec"], seed=0))
    slots, _ = cases.make_inputs(name)
    for s in slots:
        if not s.is_src:
            s.value = torch.where(s.value == om.PAD, torch.full_like(s.value, 5), s.value)
    with torch.no_grad():
        logits, _ = om.model_forward(sd, cases.oracle_cfg(name), slots)
    assert ((logits - inc).norm() / inc.norm()).item() <= 1e-5


@pytest.mark.parametrize("name", ["text_A", "patch_B"])
def test_dropout_call_sites_match_the_reference(name):
    """Row A17: tests/golden/drop_<case>.pt is the unmodified reference in training mode with every F.dropout / DropPath
    draw taken from one seeded stream (oracle/make_golden_dropout.py).  The oracle, given the same stream through
    DROP_HOOK, reproduces logits and loss only if it draws masks of the same shapes in the same order -- i.e. applies
    dropout at the reference's call sites (adaptor/base.py:181, multihead_attention.py:335, transformer_layer.py:181,195,203 /
    :433,466,481,489 and drop-path :87,:333)."""
    from oracle.make_golden_dropout import P_ACT, P_ATTN, P_PATH, P_RES, MaskStream

    fx = torch.load(os.path.join(GOLD, f"drop_{name}.pt"), weights_only=False)
    g = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    cfg = cases.oracle_cfg(name)
    slots, target = cases.make_inputs(name)
    stream = MaskStream(fx["seed"])
    src_len = sum((s.value.shape[1] if not isinstance(s.value, dict) and s.value.dim() == 2 else 0) for s in slots if s.is_src)

    def hook(kind, x):
        if kind == "embed":
            return stream.dropout_mult(x.shape, P_RES)
        if kind == "act":
            return stream.dropout_mult(x.shape, P_ACT)
        if kind == "attn_probs":
            # Mode B encoder self-attention (fast path) had its dropout switched off in the fixture: the only square
            # score matrices over the full source length
            if cfg.mode == "B" and x.shape[1] == x.shape[2] and x.shape[1] > target.shape[1]:
                return torch.ones((), dtype=x.dtype)
            return stream.dropout_mult(x.shape, P_ATTN)
        if kind == "branch":  # dropout, then drop-path per sample on T x B x C (droppath.py:41-63, batch_axis = 1)
            m = stream.dropout_mult(x.shape, P_RES)
            keep = torch.floor((1.0 - P_PATH) + stream.uniform((1, x.shape[1], 1)))
            return m * keep / (1.0 - P_PATH)
        raise AssertionError(kind)

    om.DROP_HOOK = hook
    try:
        with torch.no_grad():
            logits, _ = om.model_forward(sd, cfg, slots)
    finally:
        om.DROP_HOOK = None
    assert stream.calls == fx["draws"], (stream.calls, fx["draws"])
    assert ((logits - fx["logits"]).norm() / fx["logits"].norm()).item() <= 1e-5
    loss = om.cross_entropy_sum(logits, target)
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))


def test_box_target_golden_from_reference():
    """Row A9 / configs[4] visual_grounding: the oracle on IMAGE + TEXT -> BOX (`<bin>` tokens, BOX slot routed to the
    text adaptor) against the unmodified reference (tests/golden/box_A.pt, oracle/make_golden_box.py); bins bit-exact."""
    from oracle.make_golden_box import make_box_inputs

    fx = torch.load(os.path.join(GOLD, "box_A.pt"), weights_only=False)
    g = torch.load(os.path.join(GOLD, "resnet_A.pt"), weights_only=False)
    sd = cases.synth_state_dict(g["spec"], seed=0)
    slots, target, coords, bins = make_box_inputs()
    assert torch.equal(bins, fx["bins"])
    with torch.no_grad():
        logits, _ = om.model_forward(sd, cases.oracle_cfg("resnet_A"), slots)
    assert tuple(logits.shape) == tuple(fx["logits"].shape) == (4, 5, 512)
    assert ((logits - fx["logits"]).norm() / fx["logits"].norm()).item() <= 1e-5
    assert abs(float(om.cross_entropy_sum(logits, target)) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))


def test_spec_augment_oracle_matches_reference_transform():
    """oracle_audio.spec_augment against the output of the reference's own SpecAugmentTransform under the same numpy seed
    (tests/golden/specaugment.pt, oracle/make_golden_specaugment.py): bit-exact (integer draws + assignments)."""
    import numpy as np
    from oracle import oracle_audio as oa
    from oracle.make_golden_specaugment import CASE, case_input

    fx = torch.load(os.path.join(GOLD, "specaugment.pt"), weights_only=False)
    x = case_input()
    for tag, mv in (("zero", 0.0), ("mean", None)):
        np.random.seed(7)
        assert torch.equal(torch.from_numpy(oa.spec_augment(x, mask_value=mv, **CASE)), fx[tag]), tag
