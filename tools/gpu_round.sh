#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_incremental_gpu.py -m gpu -q -x > gpurun_out/pytest_inc.log 2>&1; echo "pytest inc rc=$?"; tail -25 gpurun_out/pytest_inc.log | cut -c1-220
