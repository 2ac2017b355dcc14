#!/bin/bash
# usage: gpu_ncu.sh <kernel-regex> <skip> <count> [bench args...]   -- plain run first, then one ncu --set full capture
mkdir -p gpurun_out
K=$1; S=$2; C=$3; shift 3
CMD="python bench.py --steps 2 --warmup 3 --batches-per-step 8 --cpu-batches 1 $@"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$K $CMD > gpurun_out/ncu_$K.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$K.log; tail -c 600 gpurun_out/plain.log
