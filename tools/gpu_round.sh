# One GPU session: parity tests, micro-benchmarks, bench lines.  Usage: gpurun -- 'bash tools/gpu_round.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "layer_norm or ln_res or res_ln or gelu" > gpurun_out/pytest_ln.log 2>&1; LN_RC=$?; echo "pytest LN rc=$LN_RC"; tail -5 gpurun_out/pytest_ln.log
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -5 gpurun_out/pytest_gemm.log
if [ $LN_RC -ne 0 ]; then echo "LN tests failed; skipping the rest"; exit 1; fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/ln_bench.py > gpurun_out/ln_bench.log 2>&1; cat gpurun_out/ln_bench.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
timeout 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench rc=$?"
cat gpurun_out/bench_b32.json
for b in 64 128; do
  timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "bench b$b rc=$?"
  cat gpurun_out/bench_b$b.json
done
