set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_r1s2.json 2> gpurun_out/bench_r1s2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r1s2.json
for k in gemm ln attn; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|ln_bwd|ln_fwd|attn_" -c 9 -o gpurun_out/prof_$k -f python tools/one_kernel.py $k > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
