#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -2 gpurun_out/bench_b64.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches_v9.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --batch 64 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
tail -2 gpurun_out/launches_v9.csv | cut -c1-300
