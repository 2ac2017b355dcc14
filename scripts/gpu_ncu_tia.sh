#!/bin/bash
# GPU-box script: the new tia test, then one ncu --set full capture of the time_interval_aware launch of the sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sampler.py -m gpu -x -q -k tia 2>&1 | tail -3
CMD="python bench.py --workload sampler_sweep --steps 1 --warmup 3 --cpu-queries 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sample_random -s 7 -c 1 -f -o gpurun_out/prof_tia $CMD > gpurun_out/ncu_tia.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tia.log
