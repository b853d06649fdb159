#!/bin/bash
# LayerNorm family in isolation at the headline shapes (rows = 64 x 265 and 64 x 64) + one full ncu capture of the GELU kernels
mkdir -p gpurun_out
timeout -s KILL 300 python tools/ln_bench.py > gpurun_out/r02_ln_bench.txt 2>&1; echo "ln rc=$?"; cat gpurun_out/r02_ln_bench.txt | cut -c1-200
LN_ROWS=16960 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_kernel|ln_bwd_kernel|ln_res_ln" -s 6 -c 6 -o gpurun_out/r02_ncu_ln python tools/one_kernel.py ln > gpurun_out/ncu_ln.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ln.log
