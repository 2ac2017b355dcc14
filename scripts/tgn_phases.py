#!/usr/bin/env python
"""Per-phase time of the one-launch TGN step (dyg_tgn_step) on the tgn_reddit config: globaltimer stamps taken by CTA 0 after
every grid barrier, averaged over the batches after a warm-up.  GPU box only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import TGNWL, REF_BATCH   # noqa: E402

wl = TGNWL()
dev = torch.device('cuda', 0)
wl.build(dev)
m = wl.model
m.phase_ns = torch.zeros(16, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
with torch.no_grad():
    for b in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
        hs = wl.stream.rows([b])
        args = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in hs)
        if len(sys.argv) > 2:
            flush.zero_()
        wl.step(*args)
        torch.cuda.synchronize()
        st = m.phase_ns.cpu().numpy()
        if b >= 20:
            rows.append(st.copy())
r = np.array(rows, dtype=np.float64)
names = ['P0 search+feat | persist+elect', 'P1 qk GEMM | messages', 'P2 cell tiles | attention', 'P3 o GEMM | commit', 'P4 LayerNorm',
         'P5 merge fc1', 'P6 merge fc2 | folded predictor fc1', '(P7 fused into P6)', 'P8 scores']
ends = [1, 2, 3, 4, 5, 6, 7, 8, 15]
prev = r[:, 0]
print(f'batches {len(rows)}  L2 flush between steps: {len(sys.argv) > 2}')
for n, e in zip(names, ends):
    d = (r[:, e] - prev) / 1e3
    print(f'{n:34s} {d.mean():7.2f} us  (min {d.min():6.2f}, max {d.max():6.2f})')
    prev = r[:, e]
print(f'{"total":34s} {((r[:, 15] - r[:, 0]) / 1e3).mean():7.2f} us')
print('P3 tile 0 of team 0: start +%.2f us after the P2 barrier, K loop %.2f us, epilogue %.2f us, then %.2f us until the P3 barrier opens' % (
    ((r[:, 9] - r[:, 3]) / 1e3).mean(), ((r[:, 10] - r[:, 9]) / 1e3).mean(), ((r[:, 11] - r[:, 10]) / 1e3).mean(), ((r[:, 4] - r[:, 11]) / 1e3).mean()))
