#!/bin/bash
# GPU-box script (round 2, first pass): new tests first, then the whole GPU suite, then the default bench line + reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_memory_api.py tests/test_gpu_models.py tests/test_gpu_sampler.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -15 gpurun_out/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 --save-dir gpurun_out/bench_r02 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real; tail -c 800 gpurun_out/bench_reference.json
