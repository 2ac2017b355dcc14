#!/bin/bash
# GPU-box script: tests, every bench workload, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -15 gpurun_out/pytest.log
for wl in dygformer_wiki tgat_myket tgn_reddit dygformer_lastfm; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; tail -c 1500 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
timeout 900 python bench.py --workload sampler_sweep --steps 5 > gpurun_out/bench_sampler.json 2> gpurun_out/bench_sampler.err; echo "sweep rc=$?"; cat gpurun_out/bench_sampler.json; tail -3 gpurun_out/bench_sampler.err
CMD="python bench.py --steps 2 --warmup 3 --batches-per-step 8 --cpu-batches 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 40 -c 4 -o gpurun_out/prof_linear $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log
