#!/bin/bash
# round 2, call 1: do the full-size BASELINE workloads run at all? (asr, cotrain, large) + attention baseline numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 900 python workloads.py asr cotrain large > gpurun_out/r02_probe_workloads.jsonl 2> gpurun_out/r02_probe_workloads.err; echo "workloads rc=$?"
cat gpurun_out/r02_probe_workloads.jsonl | cut -c1-700
tail -5 gpurun_out/r02_probe_workloads.err | cut -c1-300
timeout 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_base.txt 2>&1; echo "attn rc=$?"; cat gpurun_out/r02_attn_bench_base.txt
