#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu.log | cut -c1-300
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_tc.txt 2>&1; echo "attn rc=$?"; cat gpurun_out/r02_attn_bench_tc.txt | cut -c1-200
timeout -s KILL 600 python bench.py --no-cpu --kprofile > gpurun_out/r02_bench_tc1.json 2> gpurun_out/r02_bench_tc1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02_bench_tc1.json; tail -3 gpurun_out/r02_bench_tc1.err
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_tc1.json
timeout -s KILL 900 python workloads.py asr cotrain large > gpurun_out/r02_workloads_tc1.jsonl 2> gpurun_out/r02_workloads_tc1.err; echo "workloads rc=$?"; cut -c1-600 gpurun_out/r02_workloads_tc1.jsonl; tail -3 gpurun_out/r02_workloads_tc1.err
