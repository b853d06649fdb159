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


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b = torch.nn.Linear(6, 5), torch.nn.Linear(5, 3)
        self.unused = torch.nn.Linear(4, 4)  # an adaptor no batch touches

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _wrapper_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofasys_b200.distributed import DataParallelModel

    torch.manual_seed(0)
    net = _Net()
    model = DataParallelModel(net, bucket_bytes=64)
    assert model.unused is net.unused and hasattr(model, "no_sync") and hasattr(model, "all_reduce_grads")  # attribute forwarding
    g = torch.Generator().manual_seed(11)
    x = torch.randn(world, 2, 4, 6, generator=g)  # [rank, micro-batch, rows, features]
    ntok = 0
    for i in range(2):  # delayed-update loop (trainer.py:766-784): the exchange happens once, after the last micro-batch
        with model.no_sync() if i == 0 else __import__("contextlib").nullcontext():
            loss = model(x[rank, i]).pow(2).sum()
            loss.backward()
            model.all_reduce_grads()
        ntok += x[rank, i].shape[0]
    total = torch.tensor([float(ntok)])
    dist.all_reduce(total)
    for p in net.parameters():  # multiply_grads(world / sum(sample_size)) (trainer.py:857-860)
        if p.grad is not None:
            p.grad.mul_(world / total.item())
    if rank == 0:
        torch.save([p.grad.clone() for p in net.parameters()], out)
    dist.destroy_process_group()


def test_wrapper_equals_single_process_on_concatenated_batch(tmp_path):
    """N-rank gradients after no_sync() accumulation + all_reduce_grads() + the trainer's rescale == the gradient of ONE
    process on the concatenated batch divided by the total sample size (SURVEY 8d gate; reference trainer.py:857-860)."""
    out = str(tmp_path / "w0.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_wrapper_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    torch.manual_seed(0)
    net = _Net()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 2, 4, 6, generator=g)
    net(x.reshape(-1, 6)).pow(2).sum().backward()
    for p, got in zip(net.parameters(), res):
        want = torch.zeros_like(p) if p.grad is None else p.grad / 16.0
        assert torch.allclose(got, want, atol=1e-6), (got, want)
