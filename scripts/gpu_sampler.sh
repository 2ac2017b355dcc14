#!/bin/bash
# GPU-box script: sampler parity tests, then the sampler sweep (BASELINE configs[4]).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --workload sampler_sweep --steps 5 --cpu-queries ${1:-2000} > gpurun_out/bench_sampler.json 2> gpurun_out/bench_sampler.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_sampler.json').read().strip().splitlines()[-1])
for k, v in d['strategies'].items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
print('bit exact', d.get('parity_recent_bit_exact'), 'build_s', d['config']['csr_build_s'])
PY
tail -3 gpurun_out/bench_sampler.err
