#!/bin/bash
# GPU-box script: plain sampler sweep, then one ncu --set full capture of the sampling kernels (1 launch per strategy after warm-up).
mkdir -p gpurun_out
CMD="python bench.py --workload sampler_sweep --steps 1 --warmup 3 --cpu-queries 0"
$CMD > gpurun_out/plain_sampler.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${KERN:-sample_recent|sample_random}" -f -o gpurun_out/prof_sampler $CMD > gpurun_out/ncu_sampler.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sampler.log; tail -c 1500 gpurun_out/plain_sampler.log
