#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ctc_gpu.py -m gpu -q -x > gpurun_out/pytest_ctc.log 2>&1; echo "pytest ctc rc=$?"; tail -15 gpurun_out/pytest_ctc.log | cut -c1-200
timeout 300 python tools/bench_workloads.py asr 32 > gpurun_out/bench_asr.json 2> gpurun_out/bench_asr.err; echo "asr rc=$?"; cat gpurun_out/bench_asr.json; tail -3 gpurun_out/bench_asr.err
timeout 300 python tools/bench_workloads.py caption_resnet 32 > gpurun_out/bench_resnet.json 2> gpurun_out/bench_resnet.err; echo "resnet rc=$?"; cat gpurun_out/bench_resnet.json; tail -3 gpurun_out/bench_resnet.err
