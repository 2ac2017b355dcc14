#!/bin/bash
# GPU-box script: sampler tests; sweep with 8 / 4 lanes per query in the random-strategy kernel; ncu capture of the sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q 2>&1 | tail -5
for L in 8 4; do
DYG_RANDOM_LANES=$L timeout 600 python bench.py --workload sampler_sweep --steps 30 --cpu-queries 0 2> gpurun_out/exp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('random lanes $L', {k:(round(v['ms_per_launch'],3), round(v['ms_per_launch_median'],3), round(v['frac_of_hbm_peak'],3)) for k,v in d['strategies'].items()}, d['clocks'])"
done
tail -3 gpurun_out/exp.err
bash scripts/gpu_ncu_sampler.sh
