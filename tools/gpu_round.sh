#!/bin/bash
set -x
mkdir -p gpurun_out
for v in base f4b3 f3b3 f3b2; do
  echo "=== variant $v"
  OFAB_LIB=$PWD/ofasys_b200/variants/libofab_$v.so timeout 200 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench_$v.log | head -6
done
OFAB_LIB=$PWD/ofasys_b200/variants/libofab_f4b3.so timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_dropout_gpu.py -m gpu -x -q -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -3 gpurun_out/pytest_attn.log
timeout 300 python -m pytest tests/test_optim_gpu.py -m gpu -x -q > gpurun_out/pytest_optim.log 2>&1; echo "pytest optim rc=$?"; tail -3 gpurun_out/pytest_optim.log
