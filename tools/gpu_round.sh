# One GPU session: parity tests, micro-benchmarks, bench lines.  Usage: gpurun -- 'bash tools/gpu_round.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
# 4-CTA cluster / multicast GEMM first, alone and under a short timeout (a protocol bug would hang)
timeout 240 python -m pytest tests/test_ops_gpu.py -x -q -k "forced_tile and (128-2-4 or 256-2-4)" > gpurun_out/pytest_cl4.log 2>&1; CL_RC=$?; echo "pytest cl4 rc=$CL_RC"; tail -5 gpurun_out/pytest_cl4.log
if [ $CL_RC -ne 0 ]; then export OFAB_GEMM_CL=2; echo "CL4 FAILED -> running the rest with OFAB_GEMM_CL=2"; fi
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm and not (128-2-4 or 256-2-4)" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -5 gpurun_out/pytest_gemm.log
( time timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ops_gpu.py::test_gemm_forced_tile_configs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
OFAB_GEMM_CL=2 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench_cl2.log 2>&1; cat gpurun_out/gemm_bench_cl2.log
timeout 600 python bench.py --steps 20 --warmup 5 --kprofile --no-cpu > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?"
cat gpurun_out/bench_b64.json; tail -3 gpurun_out/bench_b64.err
OFAB_GEMM_CL=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_b64_cl2.json 2> gpurun_out/bench_b64_cl2.err; echo "bench cl2 rc=$?"
cat gpurun_out/bench_b64_cl2.json
