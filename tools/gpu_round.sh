# One GPU session: parity tests, micro-benchmarks, bench lines.  Usage: gpurun -- 'bash tools/gpu_round.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 400 python -m pytest tests/test_ops_gpu.py -q -k "splitk" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest splitk rc=$?"; tail -3 gpurun_out/pytest_gemm.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -3 gpurun_out/bench_b64.err
timeout 600 python bench.py --steps 20 --warmup 5 --batch 32 --no-cpu > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench b32 rc=$?"
cat gpurun_out/bench_b32.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/bench_ref.json
