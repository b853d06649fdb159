#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "chunked or cross_entropy or fused" > gpurun_out/r02_pytest_ce.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_ce.log | cut -c1-300
timeout -s KILL 600 python bench.py --no-cpu --no-workloads > gpurun_out/r02_bench_ce_oneshot.json 2> gpurun_out/r02_bench_ce_oneshot.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ce_oneshot.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['peak_mem_gb'], d['ce_chunk_rows'], d['clocks'])"
OFAB_CE_CHUNK_ROWS=1024 timeout -s KILL 600 python bench.py --no-cpu --no-workloads > gpurun_out/r02_bench_ce_chunk1024.json 2> gpurun_out/r02_bench_ce_chunk1024.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ce_chunk1024.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['peak_mem_gb'], d['ce_chunk_rows'], d['clocks'])"
OFAB_CE_CHUNK_ROWS=512 timeout -s KILL 600 python bench.py --no-cpu --no-workloads > gpurun_out/r02_bench_ce_chunk512.json 2> gpurun_out/r02_bench_ce_chunk512.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_ce_chunk512.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['peak_mem_gb'], d['ce_chunk_rows'], d['clocks'])"
timeout -s KILL 600 python workloads.py asr --kprofile > gpurun_out/r02_workloads_asr.jsonl 2> gpurun_out/r02_workloads_asr.err; echo "asr rc=$?"; cut -c1-300 gpurun_out/r02_workloads_asr.jsonl
