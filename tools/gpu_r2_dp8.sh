#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index --format=csv,noheader | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613"
date +%s
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout -s KILL 420 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --workloads asr --kprofile > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "n8 rc=$?"; date +%s
cp gpurun_out/kprofile.json gpurun_out/r02_kprofile_n8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['dp_check'])
for w in d['workloads']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k in ('workload','ms_per_step','seq_per_s','n_gpus','error','cuda_graph')})
PY
grep -i "NVLS\|capture failed" gpurun_out/r02_bench_n8.err | head -5
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-workloads > gpurun_out/r02_bench_n1_samebox.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_n1_samebox.json
