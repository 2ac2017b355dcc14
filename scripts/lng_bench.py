import sys, torch
sys.path.insert(0, '.')
from dyglib_b200 import ops
M, D, N = 819200, 200, 816
torch.manual_seed(0)
x = torch.randn(M, D, device='cuda'); w = torch.randn(N, D, device='cuda') / 14; b = torch.randn(N, device='cuda')
g = torch.ones(D, device='cuda'); be = torch.zeros(D, device='cuda')
ws = ops.split_bf16(w)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl = ops.ln_gemm(x, g, be, 1e-5, ws, b, want='split'); e1.record(); torch.cuda.synchronize()
    print('ln_gemm ms', e0.elapsed_time(e1), flush=True)
