#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_optim_gpu.py -m gpu -x -q > gpurun_out/pytest_optim.log 2>&1; echo "pytest optim rc=$?"; tail -12 gpurun_out/pytest_optim.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"adam_" -c 6 -o gpurun_out/prof_adam -f python tools/one_kernel.py adam > gpurun_out/ncu_adam.log 2>&1; echo "ncu adam rc=$?"
timeout 200 ncu -i gpurun_out/prof_adam.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/ncu_adam_raw.csv 2>/dev/null; head -c 1500 gpurun_out/ncu_adam_raw.csv
