import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dyglib_b200 import ops
torch.manual_seed(0)
P, B, ns, nd = 2, 3200, 32, 32
F, T, C = 172, 100, 50
N, E, maxc = 9228, 157475, 128
dev = 'cuda'
node = torch.randn(N, F, device=dev); edge = torch.randn(E, F, device=dev); lut = torch.randn(maxc + 1, C, device=dev)
tw = (1.0 / 10 ** torch.linspace(0, 9, T, device=dev)).float().contiguous(); tb = torch.zeros(T, device=dev)
ws = [torch.randn(C, P * w, device=dev) for w in (F, F, T, C)]
bias = torch.randn(4 * C, device=dev)
tq = 2e5 + torch.rand(B, device=dev, dtype=torch.float64) * 1e5
S = ns + nd
X = torch.empty(B * S, 4 * C, device=dev)
sides = []
for ntok, off in ((ns, 0), (nd, ns)):
    Lp = ntok * P
    sides.append((torch.randint(1, N, (B, Lp), device=dev), torch.randint(1, E, (B, Lp), device=dev), (torch.rand(B, Lp, device=dev) * 2e5).float(),
                  torch.randint(0, maxc, (B, Lp), device=dev), torch.randint(0, maxc, (B, Lp), device=dev), ntok, off))
packed = ops.pack_patch_weights(*ws, P)
npl, epl, lpl = ops.table_planes(node), ops.table_planes(edge), ops.table_planes(lut)
def run():
    ops.patch_project(sides, npl, F, epl, F, lpl, C, tq, tw, tb, packed, bias, P, C, S, X)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print('DBG', os.environ.get('DYG_PP_DBG'), 'tokens', B * S, 'ms/launch', e0.elapsed_time(e1) / 5)
