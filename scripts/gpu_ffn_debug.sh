#!/bin/bash
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/ffn_first.log 2>&1 <<'PY'
import torch, numpy as np
from dyglib_b200 import ops
torch.manual_seed(0)
for (M, D, Dff) in [(256, 200, 32), (256, 200, 64), (256, 200, 800), (1000, 200, 800), (100000, 200, 800)]:
    x = torch.randn(M, D, device='cuda'); gm = torch.ones(D, device='cuda'); bt = torch.zeros(D, device='cuda')
    w1 = torch.randn(Dff, D, device='cuda') / np.sqrt(D); b1 = torch.randn(Dff, device='cuda')
    w2 = torch.randn(D, Dff, device='cuda') / np.sqrt(Dff); b2 = torch.randn(D, device='cuda')
    got = ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2)
    torch.cuda.synchronize()
    xd = x.double(); y = torch.nn.functional.layer_norm(xd, (D,), gm.double(), bt.double(), 1e-5)
    want = xd + torch.nn.functional.gelu(y @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double()
    print(M, D, Dff, 'rel err', float((got.double() - want).abs().max() / want.abs().max()), flush=True)
M = 204800
x = torch.randn(M, 200, device='cuda'); gm = torch.ones(200, device='cuda'); bt = torch.zeros(200, device='cuda')
w1 = torch.randn(800, 200, device='cuda') / 14; b1 = torch.randn(800, device='cuda'); w2 = torch.randn(200, 800, device='cuda') / 28; b2 = torch.randn(200, device='cuda')
out = torch.empty_like(x)
for _ in range(3): ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2, out=out)
e1.record(); torch.cuda.synchronize()
print('ln_ffn M=204800 us', e0.elapsed_time(e1) / 5 * 1e3)
PY
echo "first rc=$?"; tail -12 gpurun_out/ffn_first.log
