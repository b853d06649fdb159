# One GPU session: parity tests, micro-benchmarks, bench lines.  Usage: gpurun -- 'bash tools/gpu_round.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 400 python tools/gemm_sweep.py > gpurun_out/gemm_sweep.log 2>&1; cat gpurun_out/gemm_sweep.log
for k in gemm ln attn; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|ln_bwd|ln_fwd|attn_" -c 9 -o gpurun_out/prof_$k -f python tools/one_kernel.py $k > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu --batch 64 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
tail -3 gpurun_out/launches.csv | cut -c1-300
