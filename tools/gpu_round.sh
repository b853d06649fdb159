#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench_idx16.log 2>&1; cat gpurun_out/attn_bench_idx16.log
timeout 300 python tools/bench_workloads.py asr 32 > gpurun_out/bench_asr.json 2> gpurun_out/bench_asr.err; echo "asr rc=$?"; cat gpurun_out/bench_asr.json
