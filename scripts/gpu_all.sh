#!/bin/bash
# GPU-box script: all GPU tests, every bench workload (ours + reference arm for the headline), sampler sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for wl in dygformer_wiki tgat_myket tgn_reddit dygformer_lastfm tgat_train; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; tail -c 1200 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_reference.json
timeout 900 python bench.py --workload sampler_sweep --steps 5 > gpurun_out/bench_sampler.json 2> gpurun_out/bench_sampler.err; echo "sweep rc=$?"; cat gpurun_out/bench_sampler.json; tail -3 gpurun_out/bench_sampler.err
