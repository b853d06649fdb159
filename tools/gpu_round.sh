#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/flaky_probe.py text_A 40 > gpurun_out/flaky_pdl1.log 2>&1; tail -12 gpurun_out/flaky_pdl1.log
OFAB_PDL=0 timeout 300 python tools/flaky_probe.py text_A 40 > gpurun_out/flaky_pdl0.log 2>&1; tail -12 gpurun_out/flaky_pdl0.log
timeout 300 python -m pytest tests/test_audio_gpu.py -m gpu -q > gpurun_out/pytest_audio.log 2>&1; echo "pytest audio rc=$?"; tail -5 gpurun_out/pytest_audio.log
