import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dyglib_b200 import ops
torch.manual_seed(0)
M = int(os.environ.get('M', 204800))
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def bench(name, N, K, **kw):
    a = ops.split_bf16(torch.randn(M, K, device=dev))
    w = torch.randn(N, K, device=dev) / np.sqrt(K)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if kw.pop('res', False) else None
    want = kw.pop('want', 'f32')
    for _ in range(2): ops.gemm(a, w, b, residual=r, want=want, **kw)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, b, residual=r, want=want, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f'{name:6s} N={N:4d} K={K:4d} {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:7.1f} TFLOP/s(fp32-equiv)', flush=True)
bench('qkv', 600, 200)
bench('out', 200, 200, res=True)
bench('ffn1', 800, 200, act=ops.ACT_GELU, want='split')
bench('ffn1n', 800, 200, want='split')
bench('ffn2', 200, 800, res=True)
bench('outnr', 200, 200)
bench('outsp', 200, 200, want='split')
bench('k32', 200, 32)
bench('n208', 208, 208)
