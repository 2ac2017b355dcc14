"""Per-launch timing of sample_recent on the sweep graph (no clock sampling): is the run-to-run spread per process or per launch?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dyglib_b200.utils.utils import NeighborSampler

dev = torch.device('cuda', 0)
E, Q, k = 100_000_000, 1 << 24, 20
src, dst, eid, t, num_nodes = bench.device_power_law_graph(E, int(E * 0.08), int(E * 0.02), 5, dev)
s = NeighborSampler(None, 'recent', 0.0, 0, dev, 'philox', 'device', _edges=(src, dst, eid, t, num_nodes, True))
gen = torch.Generator(device=dev).manual_seed(1234)
qe = torch.randint(int(E * 0.7), E, (Q,), generator=gen, device=dev)
coin = torch.rand(Q, generator=gen, device=dev) < 0.5
nodes = torch.where(coin, src[qe], dst[qe]).contiguous()
times = t[qe].contiguous()
del src, dst, eid, qe, coin
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(40):
        out = s.get_historical_neighbors_device(nodes, times, k)
        ev[i + 1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(40)]
    print(f'rep {rep}: host enqueue {1e3 * (t1 - t0):.1f} ms; per-launch ms min {min(ms):.3f} median {sorted(ms)[20]:.3f} max {max(ms):.3f}; total {ev[0].elapsed_time(ev[40]):.1f}')
    print('  ', ' '.join(f'{x:.2f}' for x in ms))
    time.sleep(0.5)
