#!/bin/bash
# GPU-box script: first runs of the TMA/tcgen05 GEMM, each step under its own timeout so a hang cannot eat the lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/gpu.txt
timeout 120 python - > gpurun_out/gemm_first.log 2>&1 <<'PY'
import torch, numpy as np
from dyglib_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(128, 16, 32), (128, 208, 32), (128, 200, 200), (1000, 600, 200), (640, 200, 800)]:
    a = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / np.sqrt(K)
    sa = ops.split_bf16(a)
    got = ops.gemm(sa, w)
    torch.cuda.synchronize()
    want = a.double() @ w.double().t()
    err = float((got.double() - want).abs().max() / want.abs().max())
    print(M, N, K, 'rel err', err, flush=True)
PY
echo "first rc=$?"; tail -20 gpurun_out/gemm_first.log
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/pytest_gemm.log 2>&1; echo "gemm rc=$?"; tail -25 gpurun_out/pytest_gemm.log
