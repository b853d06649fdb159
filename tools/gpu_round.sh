# One GPU session: parity tests, micro-benchmarks, bench lines.  Usage: gpurun -- 'bash tools/gpu_round.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; G_RC=$?; echo "pytest gemm rc=$G_RC"; tail -5 gpurun_out/pytest_gemm.log

timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile --no-cpu > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -3 gpurun_out/bench_b64.err
