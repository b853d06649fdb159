"""CPU, world_size 2, gloo: the data-parallel exchange step (gradient average over ranks, bucketed)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofasys_b200.distributed import allreduce_grads, build_buckets

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (5, 300, 70000, 17, 4096)]
    for i, p in enumerate(params):
        if i == 3 and rank == 1:
            continue  # an adaptor this rank's task did not touch: contributes zeros
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1)
    buckets = build_buckets(params, bucket_bytes=1 << 16)
    assert len(buckets) >= 2
    allreduce_grads(buckets, scale=2.0)
    res = [p.grad.clone() for p in params]
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_bucketed_gradient_average(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    for i, g in enumerate(res):
        expect = ((1 + 2) / 2 if i != 3 else (1 + 0) / 2) * (i + 1) * 2.0
        assert torch.allclose(g, torch.full_like(g, expect)), (i, g[:3], expect)
