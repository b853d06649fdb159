"""torchrun worker for tests/test_ops_gpu.py::test_dp_nccl_two_gpus: every rank holds different gradients; after the
flat-bucket exchange every rank must hold the mean."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ofasys_b200.distributed import GradBuckets  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
shapes = [(768,), (3, 5, 7), (1000, 768), (1,), (257, 33)]
params = [torch.nn.Parameter(torch.zeros(s, dtype=torch.bfloat16, device=dev)) for s in shapes]
per_rank = []
for r in range(world):
    g = torch.Generator().manual_seed(100 + r)
    per_rank.append([torch.randn(s, generator=g).bfloat16() for s in shapes])
for p, t in zip(params, per_rank[rank]):
    p.grad = t.to(dev)
params[3].grad = None if rank == 1 else params[3].grad  # an adaptor this rank's task did not touch
gb = GradBuckets(params, bucket_bytes=1 << 20)
gb.allreduce()
torch.cuda.synchronize()
for i, p in enumerate(params):
    terms = [per_rank[r][i].float() if not (i == 3 and r == 1) else torch.zeros(shapes[i]) for r in range(world)]
    want = sum(terms) / world
    err = (p.grad.float().cpu() - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), (i, err)
dist.barrier()
if rank == 0:
    print("DP_NCCL_OK")
dist.destroy_process_group()
