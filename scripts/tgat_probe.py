import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from dyglib_b200 import ops
from dyglib_b200.utils.graph import GraphedStep
wl = bench.make_workload('tgat_myket'); dev = torch.device('cuda'); wl.build(dev)
st = wl.stream
G = 8
hs = [st.rows(bench.shard_batches(i, G, 1, 0, st.nb)) for i in range(6)]
ds = [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in h) for h in hs]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10, fl=True):
    tot = 0
    for i in range(n):
        if fl: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*ds[i % 6]); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
with torch.no_grad():
    for i in range(3): wl.step(*ds[i])
    print('direct flush', t(wl.step), 'noflush', t(wl.step, fl=False))
    g = GraphedStep(wl.step, ds[0], warmup=2)
    for i in range(3): g(*ds[i])
    print('graph flush', t(g), 'noflush', t(g, fl=False))
    n0 = ops.launch_count; wl.step(*ds[0]); print('launches', ops.launch_count - n0)
    # per-phase: time each op via PROFILE
    ops.PROFILE = []
    flush.zero_(); wl.step(*ds[1]); torch.cuda.synchronize()
    for k, e0, e1, fl, by in ops.PROFILE: print(f'{k:32s} {e0.elapsed_time(e1)*1000:8.1f} us')
