#!/bin/bash
# 2 GPUs: data-parallel gates + arena vs split exchange
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "dp_ or arena or layerwise" > gpurun_out/r02_pytest_dp2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_dp2.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout -s KILL 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-workloads --dp arena > gpurun_out/r02_bench_n2_arena.json 2> gpurun_out/r02_bench_n2_arena.err; echo "arena rc=$?"; cut -c1-260 gpurun_out/r02_bench_n2_arena.json; grep -i "capture\|error" gpurun_out/r02_bench_n2_arena.err | head -5
timeout -s KILL 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-workloads --dp split > gpurun_out/r02_bench_n2_split.json 2> gpurun_out/r02_bench_n2_split.err; echo "split rc=$?"; cut -c1-260 gpurun_out/r02_bench_n2_split.json
NCCL_PROTO=Simple timeout -s KILL 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-workloads --dp arena > gpurun_out/r02_bench_n2_arena_simple.json 2> gpurun_out/r02_bench_n2_arena_simple.err; echo "arena simple rc=$?"; cut -c1-260 gpurun_out/r02_bench_n2_arena_simple.json
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-workloads > gpurun_out/r02_bench_n1_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_n1_ref.json
