#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -3 gpurun_out/bench_b64.err
cp gpurun_out/kprofile.json gpurun_out/kprofile_pdl.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-pdl --no-cpu > gpurun_out/bench_b64_nopdl.json 2> gpurun_out/bench_b64_nopdl.err; echo "bench nopdl rc=$?"
cat gpurun_out/bench_b64_nopdl.json; tail -3 gpurun_out/bench_b64_nopdl.err
