import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dyglib_b200 import ops
M = int(os.environ.get("FFN_M", 204800))
x = torch.randn(M, 200, device='cuda'); gm = torch.ones(200, device='cuda'); bt = torch.zeros(200, device='cuda')
w1 = torch.randn(800, 200, device='cuda') / 14; b1 = torch.randn(800, device='cuda'); w2 = torch.randn(200, 800, device='cuda') / 28; b2 = torch.randn(200, device='cuda')
out = torch.empty_like(x)
for _ in range(3): ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2, out=out)
e1.record(); torch.cuda.synchronize()
print('DBG', os.environ.get('DYG_FFN_DBG'), 'ln_ffn M=204800 us', e0.elapsed_time(e1) / 5 * 1e3)
