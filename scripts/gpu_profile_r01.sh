#!/bin/bash
# GPU-box script: round-1 profile set.  Every ncu pass runs only after the same command exited 0 without ncu.
# Reports are summarised on the box (gpurun_out/ is capped at 64 MiB): raw-page CSV + the compact table; the .ncu-rep of the
# dominant kernels are kept, the rest are deleted.
mkdir -p gpurun_out
sum() { python scripts/ncu_summary.py gpurun_out/$1.ncu-rep gpurun_out/$1.md > /dev/null 2>&1; ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; }
if [ -z "$SKIP_LIST" ]; then
timeout 600 python bench.py > gpurun_out/plain_default.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_default.csv python bench.py > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
fi
CMD="python bench.py --steps 2 --warmup 3 --cpu-batches 1"
# one timed step = 14 launches of these kernels (3 warm-up steps + the parity batch come first); DYG_KERNELS / DYG_SKIP / DYG_COUNT narrow it
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${DYG_KERNELS:-ln_ffn_bf16x3|gemm_bf16x3|patch_project|seq_attention_mma|layernorm_split}" -s ${DYG_SKIP:-46} -c ${DYG_COUNT:-14} -f -o gpurun_out/prof_dygformer $CMD > gpurun_out/ncu_dygformer.log 2>&1
echo "dygformer full rc=$?"; sum prof_dygformer; rm -f gpurun_out/prof_dygformer.ncu-rep
CMD="python bench.py --workload tgat_myket --no-graph --steps 2 --warmup 3 --cpu-batches 1"
timeout 600 $CMD > gpurun_out/plain_tgat.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"temporal_attend" -s 9 -c 3 -f -o gpurun_out/prof_tattn $CMD > gpurun_out/ncu_tattn.log 2>&1
echo "tgat full rc=$?"; sum prof_tattn
if [ -n "$SKIP_SAMPLER" ]; then du -sh gpurun_out; exit 0; fi
CMD="python bench.py --workload sampler_sweep --steps 1 --warmup 3 --cpu-queries 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_recent" -s 3 -c 1 -f -o gpurun_out/prof_sampler_recent $CMD > gpurun_out/ncu_sampler.log 2>&1
echo "sampler recent rc=$?"; sum prof_sampler_recent; rm -f gpurun_out/prof_sampler_recent.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_random" -s 3 -c 5 -f -o gpurun_out/prof_sampler_random $CMD > gpurun_out/ncu_sampler2.log 2>&1
echo "sampler random rc=$?"; sum prof_sampler_random; rm -f gpurun_out/prof_sampler_random.ncu-rep
du -sh gpurun_out
