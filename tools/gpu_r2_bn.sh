#!/bin/bash
# BatchNorm rewrite + reduction tails + multi-head attention CTAs: operator tests, ResNet / full-size parity, workload timings, headline line
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_bn.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest_bn.log | cut -c1-300
timeout -s KILL 600 python workloads.py cotrain large --kprofile > gpurun_out/r02_workloads_bn.jsonl 2> gpurun_out/r02_workloads_bn.err; echo "workloads rc=$?"; cut -c1-400 gpurun_out/r02_workloads_bn.jsonl; tail -3 gpurun_out/r02_workloads_bn.err | cut -c1-200
timeout -s KILL 600 python bench.py --no-cpu --no-workloads --kprofile > gpurun_out/r02_bench_bn.json 2> gpurun_out/r02_bench_bn.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_bn.json; tail -2 gpurun_out/r02_bench_bn.err | cut -c1-200
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_bn_headline.json
