#!/bin/bash
# TGN fused-step parity tests + the tgn_reddit bench line
timeout 600 python -m pytest tests/test_gpu_memory_api.py tests/test_gpu_models.py tests/test_gpu_e2e.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --workload tgn_reddit > gpurun_out/bench_tgn2.json 2> gpurun_out/bench_tgn2.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/bench_tgn2.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["parity_max_abs_err"], d["gpu_launches"])
P
tail -3 gpurun_out/bench_tgn2.err
