"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch, fp32) of the optimizer step that follows the measured path
every update.  Only tests/ and bench.py's checker legs may import it.

Reference flow for bf16 training (`common.bf16`, no loss scaler), paths relative to /root/reference/ofasys:
  trainer.py:857-884           optimizer.multiply_grads(world / sample_size); clip_grad_norm(clip_norm); optimizer.step()
  engine/optim/fp16_optimizer.py:104-131  bf16 grads are copied to fp32 grads of the fp32 master parameters
                           :170-172  multiply_grads(c): _multiply_factor *= c              (deferred)
                           :174-189  clip_grad_norm: grad_norm = factor * ||g||; clip_coef = clamp(max_norm / (grad_norm + 1e-6), max=1);
                                     factor *= clip_coef                                     (deferred)
                           :152-168,191-204  step: fp32 grads *= factor (ONE multiply), fp32 Adam step, masters copied to bf16
  module/utils.py:342-384      total norm = norm of the per-tensor fp32 L2 norms
  engine/optim/adam.py:150-216 Adam with decoupled weight decay (p -= wd * lr * p before the update), bias-corrected step size

Pinning: tests/golden/optim_adam.pt holds the outputs of the reference's own `Adam.step` and `clip_grad_norm_`
(imported unmodified by oracle/make_golden_optim.py) on seeded tensors; tests/test_oracle_golden.py replays them here.
"""
import math
from typing import List

import torch


def total_norm(grads32: List[torch.Tensor]) -> torch.Tensor:
    """module/utils.py:358-374 (fallback path without apex's multi_tensor_l2norm)."""
    if len(grads32) == 1:
        return torch.norm(grads32[0], p=2, dtype=torch.float32)
    return torch.norm(torch.stack([torch.norm(g, p=2, dtype=torch.float32) for g in grads32]))


def adam_step_(p32, g32, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """engine/optim/adam.py:195-213 on fp32 tensors, in place.  `step` is the 1-based update count."""
    beta1, beta2 = betas
    exp_avg.mul_(beta1).add_(g32, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g32, g32, value=1 - beta2)
    denom = exp_avg_sq.sqrt().add_(eps)
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    step_size = lr * math.sqrt(bias_correction2) / bias_correction1
    if weight_decay != 0:
        p32.add_(p32, alpha=-weight_decay * lr)
    p32.addcdiv_(exp_avg, denom, value=-step_size)


def total_norm_exact(grads32: List[torch.Tensor]) -> torch.Tensor:
    """The same quantity accumulated in float64 and rounded once.  The reference's fp32 accumulation (total_norm above,
    run on CPU) is ~1.5e-5 relative away from it on the golden case; the CUDA kernel (fp32 per-thread partials of 64
    elements, float64 across CTAs) is within 1e-8.  Tests that compare optimizer STATE to 2e-6 use this norm so the clip
    coefficient is not the dominant difference; the fixture comparison keeps the reference's own norm."""
    return torch.sqrt(sum((g.double() ** 2).sum() for g in grads32)).float()


def update(masters, grads_bf16, exp_avgs, exp_avg_sqs, step, lr, betas, eps, weight_decay, grad_scale, max_norm, norm_fn=total_norm):
    """One trainer update on lists of tensors (in place on masters / moments).
    Returns (grad_norm, new bf16 parameters).  grads_bf16[i] may be None (unused parameter -> zero gradient,
    fp16_optimizer.py:129-130)."""
    g32 = [torch.zeros_like(m) if g is None else g.float() for g, m in zip(grads_bf16, masters)]  # :104-131
    factor = float(grad_scale)  # :170-172
    grad_norm = factor * norm_fn(g32)  # :178
    if max_norm > 0.0:
        clip_coef = (max_norm / (grad_norm + 1e-6)).clamp_(max=1)  # :185-187
        factor = factor * clip_coef
    for g in g32:
        g.mul_(factor)  # :152-168 (fairseq_optimizer.py multiply_grads)
    for p, g, m, v in zip(masters, g32, exp_avgs, exp_avg_sqs):
        adam_step_(p, g, m, v, step, lr, betas, eps, weight_decay)
    return grad_norm, [p.to(torch.bfloat16) for p in masters]  # :134-150


def make_case(seed=0, shapes=((1000, 64), (77,), (3, 5, 129), (4096,), (1, 1), (513, 768)), none_at=(4,)):
    """Seeded parameters / gradient sequences shared by the golden generator and the tests."""
    g = torch.Generator().manual_seed(seed)
    params = [(torch.randn(s, generator=g) * 0.05).to(torch.bfloat16) for s in shapes]
    steps = []
    for k in range(3):
        gs = []
        for i, s in enumerate(shapes):
            if i in none_at and k == 1:
                gs.append(None)
            else:
                gs.append((torch.randn(s, generator=g) * (10.0 if k == 0 else 0.01)).to(torch.bfloat16))
        steps.append(gs)
    hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=1.0)
    scales = [1.0 / 48.0, 8.0 / 1000.0, 1.0]
    return params, steps, hyper, scales
