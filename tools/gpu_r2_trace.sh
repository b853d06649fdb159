#!/bin/bash
mkdir -p gpurun_out
for c in large_A text_A; do echo "== $c"; timeout -s KILL 200 python tools/parity_trace_layer.py $c 2>&1 | grep -v Warn | tail -16; done > gpurun_out/r02_parity_trace_layer.txt; cat gpurun_out/r02_parity_trace_layer.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "resnet or batch_norm or video or conv or audio_sub or spec_aug or image_norm or constrained or speech_to" > gpurun_out/r02_pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu6.log | cut -c1-300
timeout -s KILL 600 python workloads.py cotrain --kprofile > gpurun_out/r02_cotrain_pdl.json 2>/dev/null; cut -c1-420 gpurun_out/r02_cotrain_pdl.json
