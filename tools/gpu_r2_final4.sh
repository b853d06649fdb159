#!/bin/bash
# 1 GPU, final build (64-deep GEMM stages by default): GEMM + model parity tests, then the full suite, smoke, ncu launch list ->
# stamped GEMM traffic, default bench line
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300
timeout -s KILL 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log | cut -c1-300
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2900 -c 1000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu --no-workloads > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/launch_list_summary.py gpurun_out/r02_launches.csv 64 r02 > gpurun_out/r02_launches_summary.txt 2>&1; echo "summary rc=$?"; tail -6 gpurun_out/r02_launches_summary.txt | cut -c1-200
cp profiles/r02_gemm_traffic.json gpurun_out/r02_gemm_traffic.json
timeout -s KILL 1200 python bench.py --kprofile > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_final.json; tail -2 gpurun_out/r02_bench_final.err | cut -c1-200
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_final.json
