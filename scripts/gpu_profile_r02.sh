#!/bin/bash
# GPU-box script: round-2 profile set of the DyGFormer step.  Every ncu pass runs only after the same command exited 0 without ncu.
mkdir -p gpurun_out
sum() { python scripts/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1.md > /dev/null 2>&1; ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; }
CMD="python bench.py --only-headline --no-eager --cpu-batches 1"
timeout 600 $CMD > gpurun_out/plain_r02.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_list_r02.log 2>&1
echo "list rc=$?"
CMD="python bench.py --only-headline --no-eager --steps 2 --warmup 3 --cpu-batches 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_ffn_bf16x3|ln_gemm_bf16x3|patch_project|seq_attention_fold" -s ${DYG_SKIP:-42} -c ${DYG_COUNT:-7} -f -o gpurun_out/prof_r02_dygformer $CMD > gpurun_out/ncu_r02_dygformer.log 2>&1
echo "dygformer full rc=$?"; sum prof_r02_dygformer
for K in seq_attention_fold ln_gemm_bf16x3 ln_ffn_bf16x3; do
  ncu -i gpurun_out/prof_r02_dygformer.ncu-rep --page source --csv -k regex:$K > gpurun_out/prof_r02_${K}_source.csv 2>/dev/null
done
rm -f gpurun_out/prof_r02_dygformer.ncu-rep
du -sh gpurun_out
