#!/bin/bash
# round-end check on one B200: GPU parity suite, smoke, headline bench (with the CPU baseline leg)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -2 gpurun_out/bench_b64.err
