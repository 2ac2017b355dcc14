#!/bin/bash
# GPU-box script: whole GPU suite + smoke + default bench line (all BASELINE configs) on 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 1200 python bench.py --steps 10 --warmup 3 --save-dir gpurun_out/bench_r02 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_default.json'))
print('headline', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity_max_abs_err'))
for k, v in d.get('workloads', {}).items():
    print(k, {kk: (round(vv) if isinstance(vv, float) and vv > 10 else vv) for kk, vv in v.items() if kk in ('value', 'ms_per_step', 'parity_max_abs_err', 'error', 'gpu_launches')},
          'e2e', round(v['e2e']['value']) if 'e2e' in v else None, 'cpu', round(v['cpu_baseline']['value']) if 'cpu_baseline' in v else None,
          'eager', v.get('gpu_eager_baseline', {}).get('value'))
PY
tail -3 gpurun_out/bench_default.err
