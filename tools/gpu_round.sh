#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nvidia-smi topo -m | head -12
timeout 300 python -m pytest tests/test_cotrain_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "cotrain or box or text_only or split" > gpurun_out/pytest_cotrain.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_cotrain.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --kprofile --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
cp gpurun_out/kprofile.json gpurun_out/kprofile_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-overlap > gpurun_out/bench_n2_noov.json 2> gpurun_out/bench_n2_noov.err; echo "bench n2 noov rc=$?"
cat gpurun_out/bench_n2_noov.json
