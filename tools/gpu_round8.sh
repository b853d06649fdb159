#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
cat gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
