#!/bin/bash
# backward attention kernels walking several heads per CTA: correctness sweep, attention tests, timing tables, model parity
mkdir -p gpurun_out
timeout -s KILL 300 python tools/attn_tc_debug.py > gpurun_out/r02_attn_tc_debug6.txt 2>&1; echo "debug rc=$?"; grep -c "^ok" gpurun_out/r02_attn_tc_debug6.txt; grep -v "^ok" gpurun_out/r02_attn_tc_debug6.txt | cut -c1-250 | head
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attention" > gpurun_out/r02_pytest_attn6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_attn6.log | cut -c1-300
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_mh_bwd.txt 2>&1; echo "attn rc=$?"; cut -c1-120 gpurun_out/r02_attn_bench_mh_bwd.txt
ATTN_B=64 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_mh_bwd_b64.txt 2>&1; echo "attn64 rc=$?"; cut -c1-120 gpurun_out/r02_attn_bench_mh_bwd_b64.txt
OFAB_ATTN_HPC=1 ATTN_B=64 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_hpc1_b64.txt 2>&1; echo "attn64 hpc1 rc=$?"; cut -c1-120 gpurun_out/r02_attn_bench_hpc1_b64.txt | head -4
timeout -s KILL 900 python -m pytest tests/test_model_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_attn6b.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/r02_pytest_attn6b.log | cut -c1-300
