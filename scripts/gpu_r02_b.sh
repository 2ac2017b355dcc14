#!/bin/bash
# GPU-box script: the one-launch TGN step -- parity tests, then the tgn_reddit bench (graph replay and direct), compute-sanitizer on a small case
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_memory_api.py -m gpu -q -x -k "tgn or memory or TGN or checkpoint or sub_api" > gpurun_out/pytest_tgn.log 2>&1; echo "pytest tgn rc=$?"; tail -25 gpurun_out/pytest_tgn.log
timeout 600 python bench.py --workload tgn_reddit --steps 20 --warmup 5 > gpurun_out/bench_tgn.json 2> gpurun_out/bench_tgn.err; echo "tgn rc=$?"; cut -c1-2500 gpurun_out/bench_tgn.json; tail -5 gpurun_out/bench_tgn.err
timeout 600 python bench.py --workload tgn_reddit --steps 20 --warmup 5 --no-graph --cpu-batches 2 --no-eager > gpurun_out/bench_tgn_nograph.json 2> gpurun_out/bench_tgn_nograph.err; echo "tgn nograph rc=$?"; cut -c1-600 gpurun_out/bench_tgn_nograph.json
