#!/bin/bash
# GPU-box script: model parity tests, then the TGAT and TGN workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -8
for wl in tgat_myket tgn_reddit; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'], 3), d['kernels'], d['parity_max_abs_err'])
PY
  tail -3 gpurun_out/bench_$wl.err
done
