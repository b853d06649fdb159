#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "layer_norm or ln_ or gelu or dropout or drop" > gpurun_out/r02_pytest_ln3a.log 2>&1; echo "pytest ops rc=$?"; tail -3 gpurun_out/r02_pytest_ln3a.log | cut -c1-300
timeout -s KILL 300 python tools/ln_bench.py > gpurun_out/r02_ln_bench3.txt 2>&1; echo "ln rc=$?"; grep "enc64\|dec64" gpurun_out/r02_ln_bench3.txt | cut -c1-200
LN_ROWS=16960 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_kernel|ln_bwd_kernel|ln_res_ln" -s 6 -c 4 -o gpurun_out/r02_ncu_ln3 python tools/one_kernel.py ln > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout -s KILL 1500 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_ln3b.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/r02_pytest_ln3b.log | cut -c1-300
timeout -s KILL 600 python bench.py --no-cpu --no-workloads > gpurun_out/r02_bench_ln3.json 2> gpurun_out/r02_bench_ln3.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_ln3.json
