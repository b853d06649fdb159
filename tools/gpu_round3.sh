#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_incremental_gpu.py -m gpu -q > gpurun_out/pytest_inc.log 2>&1; echo "pytest inc rc=$?"; tail -3 gpurun_out/pytest_inc.log | cut -c1-200
