#!/bin/bash
# GPU-box script: the two ncu passes of B200_PROFILING.md on the default bench command (reduced to 8 reference batches
# per step so ~40 replays per launch stay short).  Plain run first; ncu only if it exits 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --batches-per-step 8 --cpu-batches 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16x3|patch_project|seq_attention_mma|layernorm_split" -s 30 -c 14 -f -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
