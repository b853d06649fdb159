#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --kprofile > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -2 gpurun_out/bench_b64.err
