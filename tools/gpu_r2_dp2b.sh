#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_model_gpu.py -m gpu -q -k "stagewise or layerwise" > gpurun_out/r02_pytest_stage.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_stage.log | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
date +%s
timeout -s KILL 420 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --workloads asr,cotrain > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "n2 rc=$?"; date +%s
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['grad_exchange'][:40], d['dp_check'])
for w in d['workloads']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k in ('workload','ms_per_step','seq_per_s','n_gpus','error','cuda_graph')})
PY
grep -i "capture failed\|error" gpurun_out/r02_bench_n2.err | head -5
