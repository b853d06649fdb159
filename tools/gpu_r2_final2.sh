#!/bin/bash
# 1 GPU, final build: full GPU suite, smoke, default bench line (workloads + cpu baseline), ncu launch list -> GEMM traffic stamped
# with this build, isolated attention / LayerNorm tables, full ncu captures of the attention forward and the GELU LayerNorm kernels,
# GPU-eager oracle diagnostic.
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300
timeout -s KILL 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-300
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2900 -c 1000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu --no-workloads > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/launch_list_summary.py gpurun_out/r02_launches.csv 64 r02 > gpurun_out/r02_launches_summary.txt 2>&1; echo "summary rc=$?"; tail -5 gpurun_out/r02_launches_summary.txt | cut -c1-200
cp profiles/r02_gemm_traffic.json gpurun_out/r02_gemm_traffic.json
timeout -s KILL 1200 python bench.py --kprofile > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_final.json; tail -2 gpurun_out/r02_bench_final.err | cut -c1-200
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_final.json
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_final.txt 2>&1; echo "attn rc=$?"
ATTN_B=64 timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_final_b64.txt 2>&1; echo "attn64 rc=$?"
timeout -s KILL 300 python tools/ln_bench.py > gpurun_out/r02_ln_bench_final.txt 2>&1; echo "ln rc=$?"; grep "enc64" gpurun_out/r02_ln_bench_final.txt | cut -c1-200
ATTN_B=64 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -s 2 -c 2 -o gpurun_out/r02_ncu_attn_final python tools/attn_bench.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
LN_ROWS=16960 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_kernel|ln_bwd_kernel|ln_res_ln" -s 6 -c 6 -o gpurun_out/r02_ncu_ln_final python tools/one_kernel.py ln > gpurun_out/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout -s KILL 600 python tests/bench_oracle_gpu.py > gpurun_out/r02_oracle_gpu_eager.json 2> gpurun_out/r02_oracle_gpu_eager.err; echo "oracle gpu rc=$?"; cut -c1-300 gpurun_out/r02_oracle_gpu_eager.json
