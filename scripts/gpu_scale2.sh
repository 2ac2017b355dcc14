#!/bin/bash
# N-GPU (env N, default 2) torchrun run of the default bench line (headline + workloads incl. tgat_train) and the reference arm under torchrun
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${N:-2} --steps 10 --warmup 3 > gpurun_out/bench_default_${N:-2}gpu.json 2> gpurun_out/bench_default_${N:-2}gpu.err ) 2>&1 | grep real; echo "rc=$?"
python - gpurun_out/bench_default_${N:-2}gpu.json <<'PY'
import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith('{')][-1]
d = json.loads(l)
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'n_gpus', d['n_gpus'], 'ms', round(d['ms_per_step'], 3))
for k, v in d.get('workloads', {}).items():
    print(k, round(v['value']), v.get('ms_per_step'), {kk: v[kk] for kk in ('allreduce_ms', 'ranks_in_sync', 'error') if kk in v})
PY
tail -3 gpurun_out/bench_default_${N:-2}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus ${N:-2} --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
