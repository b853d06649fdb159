#!/bin/bash
# 2 GPUs, final build: the driver's own launch line (default workloads), then the same-box N=1 line for the ratio
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout -s KILL 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['dp_check'])
for w in d['workloads']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k in ('workload','ms_per_step','seq_per_s','n_gpus','error','cuda_graph')})
PY
grep -i "capture failed\|error" gpurun_out/r02_bench_n2_final.err | head -5
timeout -s KILL 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-workloads > gpurun_out/r02_bench_n1_samebox_final.json 2>/dev/null; echo "n1 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n1_samebox_final.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
