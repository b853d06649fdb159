#!/bin/bash
mkdir -p gpurun_out
for c in text_B text_A; do timeout -s KILL 200 python tools/parity_trace.py $c 2>&1 | grep -v Warn | tail -12; done > gpurun_out/r02_parity_trace.txt; cat gpurun_out/r02_parity_trace.txt
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "arena or dp_model" > gpurun_out/r02_pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_gpu5.log | cut -c1-300
