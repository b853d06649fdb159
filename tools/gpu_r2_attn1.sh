#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/attn_tc_debug.py > gpurun_out/r02_attn_tc_debug.txt 2>&1; echo "debug rc=$?"
cat gpurun_out/r02_attn_tc_debug.txt | cut -c1-250 | tail -40
