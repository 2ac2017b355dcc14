#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q > gpurun_out/pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -15 gpurun_out/pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gemm.py > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_dygformer_wiki.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_dygformer_wiki.json; tail -3 gpurun_out/bench.err
