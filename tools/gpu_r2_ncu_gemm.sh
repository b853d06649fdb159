#!/bin/bash
mkdir -p gpurun_out
GEMM_M=16960 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/r02_ncu_gemm_final python tools/one_kernel.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; tail -2 gpurun_out/ncu_gemm.log
GEMM_B=64 timeout -s KILL 300 python tools/gemm_bench.py > gpurun_out/r02_gemm_shapes_final.txt 2>&1; echo "gemm bench rc=$?"; tail -34 gpurun_out/r02_gemm_shapes_final.txt | cut -c1-130
