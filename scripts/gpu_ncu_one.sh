#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> <skip> <name> [bench args...]  -- ncu --set full of ONE launch, source page exported on the box
mkdir -p gpurun_out
K=$1; S=$2; N=$3; shift 3
CMD="python bench.py --steps 2 --warmup 3 --cpu-batches 1 $@"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$N $CMD > gpurun_out/ncu_$N.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$N.log
ncu -i gpurun_out/prof_$N.ncu-rep --page source --csv > gpurun_out/prof_${N}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_$N.ncu-rep --page raw --csv > gpurun_out/prof_${N}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_$N*
