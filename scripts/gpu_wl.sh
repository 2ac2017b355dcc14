#!/bin/bash
# usage: gpu_wl.sh wl1 wl2 ...   -- run the GPU tests of the models and the named bench workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q > gpurun_out/pytest_models.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_models.log
for wl in "$@"; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; tail -3 gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    print('$wl', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), d['roofline']['kernel'], round(d['roofline']['frac'],4), 'parity', d.get('parity_max_abs_err'), 'launches', d['gpu_launches'], d['config'].get('launch'))
    print('   ', {k:(v['ms'],v['launches']) for k,v in d['kernels'].items()})
except Exception as e: print('$wl ERR', e)
PY
done
