"""torchrun worker for tests/test_ops_gpu.py::test_dp_model_equality: model-level data-parallel gate (workloads.dp_equality_check)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
res = workloads.dp_equality_check(dev, world, rank)
oks = [None] * world
dist.all_gather_object(oks, res["ok"])
if rank == 0:
    print("DP_MODEL " + json.dumps(res))
    print("DP_MODEL_OK" if all(oks) else "DP_MODEL_FAIL")
dist.destroy_process_group()
