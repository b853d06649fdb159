#!/bin/bash
# forward attention with several heads per CTA: correctness sweep, then graph-timed table with 1 head / CTA and with the heuristic
mkdir -p gpurun_out
timeout -s KILL 300 python tools/attn_tc_debug.py > gpurun_out/r02_attn_tc_debug5.txt 2>&1; echo "debug rc=$?"; grep -c "^ok" gpurun_out/r02_attn_tc_debug5.txt; grep -v "^ok" gpurun_out/r02_attn_tc_debug5.txt | cut -c1-250 | head
OFAB_ATTN_HPC=1 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_hpc1.txt 2>&1; echo "attn rc=$?"; cut -c1-200 gpurun_out/r02_attn_bench_hpc1.txt
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_hpcauto.txt 2>&1; echo "attn rc=$?"; cut -c1-200 gpurun_out/r02_attn_bench_hpcauto.txt
OFAB_ATTN_HPC=2 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_hpc2.txt 2>&1; echo "attn rc=$?"; cut -c1-200 gpurun_out/r02_attn_bench_hpc2.txt
ATTN_B=64 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_hpcauto_b64.txt 2>&1; echo "attn rc=$?"; cut -c1-200 gpurun_out/r02_attn_bench_hpcauto_b64.txt
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" > gpurun_out/r02_pytest_attn5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_attn5.log | cut -c1-300
