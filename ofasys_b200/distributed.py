"""Data-parallel exchange step of the hot path: the path shards by batch, every rank
holds a full replica, and the only collective is the gradient average after backward
(reference: c10d DDP wrap, ofasys/distributed/distributed_model_dispatcher.py:49-75, then
`multiply_grads(world_size / sum(sample_size))`, ofasys/engine/trainer.py:857-860).

Gradients are packed into ~32 MB flat buckets and averaged with one all-reduce per bucket
(NCCL over NVLink 5 / NVSwitch on B200; gloo in the CPU tests).  A task that did not touch an
adaptor simply contributes zeros for it (the reference needs DDP's find_unused_parameters walk).
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def build_buckets(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20) -> List[List[torch.nn.Parameter]]:
    """Reverse registration order ~ the order gradients become ready in backward."""
    buckets, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        cur.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_grads(buckets: List[List[torch.nn.Parameter]], group=None, scale: Optional[float] = None) -> None:
    """p.grad <- mean over ranks of p.grad (missing grads count as zeros); optional extra `scale`
    (the trainer's world_size / sum(sample_size))."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    for bucket in buckets:
        grads = []
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        flat = torch.cat([g.reshape(-1) for g in grads])
        if backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            if scale is not None:
                flat.mul_(scale)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_((scale or 1.0) / world)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
