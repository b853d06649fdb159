#!/bin/bash
# packed-fp32 / single-pass LayerNorm family: operator tests, model parity, isolated timings, headline line
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "layer_norm or ln_ or gelu or dropout or drop" > gpurun_out/r02_pytest_ln2a.log 2>&1; echo "pytest ops rc=$?"; tail -4 gpurun_out/r02_pytest_ln2a.log | cut -c1-300
timeout -s KILL 300 python tools/ln_bench.py > gpurun_out/r02_ln_bench2.txt 2>&1; echo "ln rc=$?"; grep "enc64\|dec64" gpurun_out/r02_ln_bench2.txt | cut -c1-200
timeout -s KILL 1500 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_ln2b.log 2>&1; echo "pytest model rc=$?"; tail -6 gpurun_out/r02_pytest_ln2b.log | cut -c1-300
timeout -s KILL 600 python bench.py --no-cpu --no-workloads > gpurun_out/r02_bench_ln2.json 2> gpurun_out/r02_bench_ln2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_ln2.json; tail -2 gpurun_out/r02_bench_ln2.err | cut -c1-200
