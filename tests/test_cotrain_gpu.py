"""GPU: the multi-task shapes of BASELINE.json configs[3] and configs[4] on the small ResNet geometry.

configs[3] "mixed-batch caption + VQA + text_infilling co-training": one model with the text and image_resnet adaptors
active takes three task batches per step and the gradients ACCUMULATE over them before the exchange (the reference's
trainer loops over the tasks of a step, engine/trainer.py:752-830); a text-only batch leaves the image adaptor unused
(zero contribution -- what DDP's find_unused_parameters handles in the reference).
configs[4] "visual_grounding (IMAGE -> BOX)": the target slot has modality BOX -- `<bin>_k` tokens produced by the
integer quantisation of preprocessor/default/box.py:101-110 (bit-exact) and routed to the text adaptor
(adaptor/general.py:36-46).

Oracle: oracle/oracle_model.py (pinned to the reference by tests/golden: resnet_A and text_A cover every function used
here).  Tolerances as in tests/test_model_gpu.py.
"""
import pytest
import torch

from oracle import cases
from oracle import oracle_model as om
from util import bf16_round_state_dict, build_product, load_golden, rel_l2, to_product_slots

pytestmark = pytest.mark.gpu
V = 512


def _tok(g, B, T, ragged=True):
    return cases._tokens(g, B, T, V, ragged)


def _target_of(prev):
    target = torch.roll(prev, -1, dims=1)
    target[:, -1] = 2
    target[prev == om.PAD] = om.PAD
    target[torch.roll(prev == om.PAD, -1, dims=1)] = om.PAD
    target[:, -1] = torch.where(prev[:, -1] == om.PAD, torch.tensor(om.PAD), torch.tensor(2))
    return target


def _tasks():
    g = torch.Generator().manual_seed(4321)
    tasks = {}
    # image_caption: image + 8-token prompt -> 12-token caption
    prev = _tok(g, 4, 12)
    prev[:, 0] = 0
    tasks["caption"] = ([om.OSlot(om.IMAGE, True, torch.randn(4, 3, 64, 64, generator=g), adaptor="image_resnet"),
                         om.OSlot(om.TEXT, True, _tok(g, 4, 8)), om.OSlot(om.TEXT, False, prev)], _target_of(prev))
    # VQA: image + 16-token question -> 8-token answer
    prev = _tok(g, 2, 8)
    prev[:, 0] = 0
    tasks["vqa"] = ([om.OSlot(om.IMAGE, True, torch.randn(2, 3, 64, 64, generator=g), adaptor="image_resnet"),
                     om.OSlot(om.TEXT, True, _tok(g, 2, 16)), om.OSlot(om.TEXT, False, prev)], _target_of(prev))
    # text_infilling: text only
    prev = _tok(g, 3, 16)
    prev[:, 0] = 0
    tasks["infill"] = ([om.OSlot(om.TEXT, True, _tok(g, 3, 24)), om.OSlot(om.TEXT, False, prev)], _target_of(prev))
    return tasks


def _setup():
    dev = torch.device("cuda:0")
    g = load_golden("resnet_A")
    sd = cases.synth_state_dict(g["spec"], seed=0)
    sd_r = bf16_round_state_dict(sd)
    cfg = cases.oracle_cfg("resnet_A")
    m = build_product("resnet_A")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    return dev, sd_r, cfg, m.to(torch.bfloat16).to(dev).train()


def _grad_errors(m, grads_ref):
    num = den = 0.0
    per = {}
    for k, p in m.named_parameters():
        gr = grads_ref[k].double()
        gp = (torch.zeros_like(p) if p.grad is None else p.grad).double().cpu()
        num += (gp - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
        per[k] = (gp, gr)
    tot = den ** 0.5
    worst = ("", 0.0)
    for k, (gp, gr) in per.items():
        if "embed_images" in k or gr.norm() <= 1e-3 * tot:  # ResNet params: see test_model_gpu (ReLU-mask flips under bf16 storage)
            continue
        e = ((gp - gr).norm() / gr.norm()).item()
        if e > worst[1]:
            worst = (k, e)
    return (num / max(den, 1e-30)) ** 0.5, worst


def test_cotraining_step_accumulates_three_tasks():
    dev, sd_r, cfg, m = _setup()
    tasks = _tasks()
    m.zero_grad(set_to_none=True)
    total = None
    for name, (slots, target) in tasks.items():
        loss_ref, _, grads = om.loss_and_grads(sd_r, cfg, slots, target)
        total = grads if total is None else {k: total[k] + grads[k] for k in total}
        loss = m.forward_loss(to_product_slots(slots, dev), target.to(dev))
        loss.backward()  # accumulates into .grad over the tasks of the step
        assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item()), (name, loss.item(), loss_ref.item())
        if name == "caption":  # integer outputs of the mixed sequence: padding mask of image tokens ++ prompt, bit-exact
            enc = m.encoder([s for s in to_product_slots(slots, dev) if s.is_src])
            enc_ref = om.encoder_forward(sd_r, cfg, [s for s in slots if s.is_src])
            assert torch.equal(enc["encoder_padding_mask"][0].cpu(), enc_ref["encoder_padding_mask"])
    torch.cuda.synchronize()
    e_grad, worst = _grad_errors(m, total)
    assert e_grad <= 3e-2, e_grad
    assert worst[1] <= 6e-2, worst


def test_text_only_batch_leaves_image_adaptor_untouched():
    """A task that never calls an adaptor contributes nothing to its parameters: .grad stays None (the exchange step
    pre-zeroes the bucket arena instead of DDP's unused-parameter graph walk)."""
    dev, sd_r, cfg, m = _setup()
    slots, target = _tasks()["infill"]
    m.zero_grad(set_to_none=True)
    m.forward_loss(to_product_slots(slots, dev), target.to(dev)).backward()
    unused = [k for k, p in m.named_parameters() if "image_resnet" in k]
    assert unused and all(dict(m.named_parameters())[k].grad is None for k in unused)
    _, _, grads = om.loss_and_grads(sd_r, cfg, slots, target)
    assert all(not grads[k].any() for k in unused)
    e_grad, worst = _grad_errors(m, grads)
    assert e_grad <= 3e-2 and worst[1] <= 6e-2, (e_grad, worst)


def test_visual_grounding_box_target():
    """IMAGE + TEXT -> BOX: 4 `<bin>` tokens (+ bos) as a BOX-modality target slot."""
    from ofasys_b200.preprocessor import quantize_box

    dev, sd_r, cfg, m = _setup()
    g = torch.Generator().manual_seed(99)
    B, num_bins, first_bin = 4, 100, V - 100
    coords = torch.rand(B, 4, generator=g) * 512
    bins = quantize_box(coords, 512, num_bins)
    assert torch.equal(bins, om.quantize_box(coords, 512, num_bins))  # integer work: bit-exact
    assert int(bins.min()) >= 0 and int(bins.max()) < num_bins
    prev = torch.cat([torch.zeros(B, 1, dtype=torch.long), first_bin + bins], dim=1)  # bos + 4 bin tokens (T = 5)
    target = torch.cat([first_bin + bins, torch.full((B, 1), 2, dtype=torch.long)], dim=1)
    slots = [om.OSlot(om.IMAGE, True, torch.randn(B, 3, 64, 64, generator=g), adaptor="image_resnet"),
             om.OSlot(om.TEXT, True, _tok(g, B, 16)), om.OSlot(om.BOX, False, prev)]
    loss_ref, logits_ref, grads = om.loss_and_grads(sd_r, cfg, slots, target)
    pslots = to_product_slots(slots, dev)
    logits, _ = m(pslots)
    assert tuple(logits.shape) == (B, 5, V)
    assert rel_l2(logits.float(), logits_ref) <= 6e-3
    m.zero_grad(set_to_none=True)
    loss = m.forward_loss(pslots, target.to(dev))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    e_grad, worst = _grad_errors(m, grads)
    assert e_grad <= 3e-2 and worst[1] <= 6e-2, (e_grad, worst)
