#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/attn_tc_debug.py > gpurun_out/r02_attn_tc_debug3.txt 2>&1; echo "debug rc=$?"; grep -c "^ok" gpurun_out/r02_attn_tc_debug3.txt; grep -v "^ok" gpurun_out/r02_attn_tc_debug3.txt | cut -c1-250 | head
timeout -s KILL 600 python tools/attn_bench.py > gpurun_out/r02_attn_bench_tc3.txt 2>&1; echo "attn rc=$?"; cat gpurun_out/r02_attn_bench_tc3.txt | cut -c1-200
ATTN_B=64 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 6 -c 9 -o gpurun_out/r02_ncu_attn_tc3 python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_attn.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu3.log | cut -c1-300
timeout -s KILL 900 python bench.py --no-cpu --kprofile > gpurun_out/r02_bench_tc3.json 2> gpurun_out/r02_bench_tc3.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_tc3.json; tail -3 gpurun_out/r02_bench_tc3.err
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_tc3.json
