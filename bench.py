#!/usr/bin/env python
"""Benchmark of the temporal neighbour-aggregation hot path (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload = BASELINE.json configs[1]: DyGFormer link prediction on the synthetic Wikipedia-shaped
graph (157,474 events, 9,228 nodes, 172-d features), patch_size 2, max_input_sequence_length 64.
A step = one pass of the hot path over `--batches-per-step` reference batches of 200 events (each batch
keeps its own padding unit): negative draw (host, precomputed like the reference's seeded sampler) ->
first-hop search + pad -> co-occurrence -> patch projections -> 2 transformer layers -> link scores for
the positive and the negative pair of every event.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dyglib_b200.synthetic import make_config_graph  # noqa: E402

REF_BATCH = 200
WORKLOADS = {
    'dygformer_wiki': dict(graph='dygformer_wiki', model='DyGFormer', P=2, L=64),
    'dygformer_lastfm': dict(graph='dygformer_lastfm', model='DyGFormer', P=16, L=512),
    'tgat_myket': dict(graph='tgat_myket', model='TGAT', k=20, layers=2),
    'tgn_reddit': dict(graph='tgn_reddit', model='TGN', k=10, layers=1),
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p['bf16_tflops_sustained'], tensor_burst=p['bf16_tflops'], source='measured')
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(',') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def dist_env():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


# ------------------------------------------------------------------------------------------------ workload data
class Stream:
    """Chronological reference batches of the evaluation region (last 15 % of events) with seeded negatives."""

    def __init__(self, g, batch=REF_BATCH, region=0.15, seed=2, start=None):
        E = g.num_interactions
        self.start = int(E * (1 - region)) if start is None else start
        self.nb = (E - self.start) // batch
        self.batch = batch
        self.g = g
        uniq = np.unique(g.dst_node_ids)
        rs = np.random.RandomState(seed)
        self.neg = uniq[rs.randint(0, len(uniq), self.nb * batch)]

    def rows(self, batch_ids):
        idx = np.concatenate([np.arange(self.start + b * self.batch, self.start + (b + 1) * self.batch) for b in batch_ids])
        nidx = np.concatenate([np.arange(b * self.batch, (b + 1) * self.batch) for b in batch_ids])
        g = self.g
        return g.src_node_ids[idx], g.dst_node_ids[idx], self.neg[nidx], g.node_interact_times[idx], g.edge_ids[idx]


def build_dygformer(g, wl, device):
    from dyglib_b200.utils.utils import get_neighbor_sampler, set_random_seed
    from dyglib_b200.models.DyGFormer import DyGFormer
    from dyglib_b200.models.modules import MergeLayer
    set_random_seed(0)
    t0 = time.perf_counter()
    sampler = get_neighbor_sampler(g, 'recent', device=device)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    model = DyGFormer(g.node_raw_features, g.edge_raw_features, sampler, 100, 50, wl['P'], 2, 2, 0.1, wl['L'], device).eval()
    pred = MergeLayer(172, 172, 172, 1).to(device).eval()
    return sampler, model, pred, build_s


def dygformer_step(model, pred, src, dst, neg, t):
    """pos and neg pairs of every event; every reference batch is its own padding unit for pos and for neg."""
    from dyglib_b200 import ops
    s2, d2, t2 = torch.cat([src, src]), torch.cat([dst, neg]), torch.cat([t, t])
    es, ed = model.compute_src_dst_node_temporal_embeddings(s2, d2, t2, batch_size=REF_BATCH)
    h = ops.linear([ops.seg_rows(es), ops.seg_rows(ed)], es.shape[0], pred.fc1.weight.detach(), pred.fc1.bias.detach(), act=ops.ACT_RELU)
    return ops.linear([ops.seg_rows(h)], h.shape[0], pred.fc2.weight.detach(), pred.fc2.bias.detach(), act=ops.ACT_SIGMOID)


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def oracle_dygformer(g, wl, state_dict, pred_sd):
    from oracle.sampler import OracleSampler
    from oracle.models import OracleDyGFormer, merge_layer
    samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
    m = OracleDyGFormer(state_dict, g.node_raw_features, g.edge_raw_features, samp, 50, wl['P'], 2, 2, wl['L'])

    def step(src, dst, neg, t):
        out = []
        with torch.no_grad():
            for d in (dst, neg):
                a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t)
                out.append(torch.sigmoid(merge_layer(pred_sd, '', a, b)))
        return out
    return step


def cpu_state_dicts(wl, g):
    """Reference-default initialised weights under seed 0, built on CPU from the package's parameter containers."""
    from dyglib_b200.utils.utils import set_random_seed
    from dyglib_b200.models.DyGFormer import DyGFormer
    from dyglib_b200.models.modules import MergeLayer
    set_random_seed(0)
    m = DyGFormer(g.node_raw_features[:2], g.edge_raw_features[:2], None, 100, 50, wl['P'], 2, 2, 0.1, wl['L'], 'cpu')
    p = MergeLayer(172, 172, 172, 1)
    return ({k: v.detach().clone() for k, v in m.state_dict().items()}, {k: v.detach().clone() for k, v in p.state_dict().items()})


def time_cpu(step, stream, n_batches, warm=1):
    for b in range(warm):
        step(*stream.rows([b])[:4])
    t0 = time.perf_counter()
    for b in range(warm, warm + n_batches):
        step(*stream.rows([b])[:4])
    return n_batches * stream.batch / (time.perf_counter() - t0)


def run_reference(args, wl_name, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    g = make_config_graph(wl['graph'])
    sd, psd = cpu_state_dicts(wl, g)
    stream = Stream(g)
    step = oracle_dygformer(g, wl, sd, psd)
    per_step = 2                                           # bounded sample: 2 reference batches per step
    for w in range(args.warmup):
        step(*stream.rows([w % stream.nb])[:4])
    times = []
    b = args.warmup
    for _ in range(args.steps):
        t0 = time.perf_counter()
        for _ in range(per_step):
            step(*stream.rows([b % stream.nb])[:4])
            b += 1
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = args.steps * per_step * REF_BATCH / total
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': 'link-pred events/sec', 'value': value, 'unit': 'events/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f"{wl_name} (DyGFormer P={wl['P']} L={wl['L']}, recent first-hop history, batch 200, pos+neg pairs)"},
        'cpu_baseline': {'value': value, 'unit': 'events/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{per_step} reference batches of 200 events per step, oracle/ (torch-CPU port of the reference path; '
                                   f'python sampler single-core, torch ops {cores} threads)'},
        'e2e': {'value': value, 'unit': 'events/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, wl_name, wl):
    rank, world, local = dist_env()
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device('cuda', torch.cuda.current_device())
    from dyglib_b200 import ops
    pk = peaks()
    g = make_config_graph(wl['graph'])
    sampler, model, pred, build_s = build_dygformer(g, wl, dev)
    stream = Stream(g)
    G = args.batches_per_step
    K, W = args.steps, args.warmup

    def step_batches(i):
        # reference batches are sharded round-robin over ranks (whole batches: the padding unit must stay intact)
        return [((i * G + j) * world + rank) % stream.nb for j in range(G)]

    host_steps = [stream.rows(step_batches(i)) for i in range(W + K)]
    dev_steps = [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in hs[:4]) for hs in host_steps]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ---------------- device-resident timing
        for i in range(W):
            dygformer_step(model, pred, *dev_steps[i])
        barrier()
        clocks = ClockSampler(torch.cuda.current_device())
        launches0 = ops.launch_count
        evs = []
        for i in range(W, W + K):
            flush.zero_()                                     # L2 flush between timed iterations (outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            scores = dygformer_step(model, pred, *dev_steps[i])
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = ops.launch_count - launches0
        clk = clocks.stop()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        # ---------------- end to end: host buffers in, host scores out, copies inside the timed region
        pinned = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in hs[:4]) for hs in host_steps]
        out_host = torch.empty((2 * G * REF_BATCH, 1), dtype=torch.float32).pin_memory()
        for i in range(W):
            sc = dygformer_step(model, pred, *[a.to(dev, non_blocking=True) for a in pinned[i]])
            out_host.copy_(sc, non_blocking=True)
        barrier()
        e2e_evs = []
        for i in range(W, W + K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sc = dygformer_step(model, pred, *[a.to(dev, non_blocking=True) for a in pinned[i]])
            out_host.copy_(sc, non_blocking=True)
            e1.record()
            e2e_evs.append((e0, e1))
        barrier()
        e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs)
        h2d = sum(a.numel() * a.element_size() for a in pinned[0])
        d2h = out_host.numel() * 4
        # ---------------- roofline pass: same steps, every launch bracketed by CUDA events on its stream
        ops.PROFILE = []
        for i in range(W, W + K):
            flush.zero_()
            dygformer_step(model, pred, *dev_steps[i])
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = tt.tolist()
        # final score gather over NCCL (outside the timed region; verifies every rank produced scores)
        gathered = torch.empty((world,) + tuple(scores.shape), device=dev, dtype=scores.dtype)
        dist.all_gather_into_tensor(gathered, scores.contiguous())
        checksum = float(gathered.double().sum().item())
    else:
        checksum = float(scores.double().sum().item())
    per_kernel = {}
    for name, e0, e1, fl, by in prof:
        d = per_kernel.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1)
        d[1] += fl
        d[2] += by
        d[3] += 1
    events_total = K * G * REF_BATCH * world
    value = events_total / (total_ms * 1e-3)
    top = max(per_kernel.items(), key=lambda kv: kv[1][0])
    name, (ms, fl, by, cnt) = top
    achieved = fl / (ms * 1e-3) / 1e12
    roofline = {'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': pk['tensor'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tensor'], 'traffic': None, 'peak_source': pk['source'] + ' (bf16 sustained)',
                'launches': cnt, 'avg_launch_us': 1e3 * ms / cnt,
                'share_of_step': ms / max(sum(v[0] for v in per_kernel.values()), 1e-9),
                'note': 'fp32 FFMA tiles in round 1 (tcgen05 path not yet wired); flops = sum 2*M*N*K over launches'}
    line = {
        'metric': 'link-pred events/sec', 'value': value, 'unit': 'events/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f"{wl_name} (DyGFormer P={wl['P']} L={wl['L']}, recent first-hop history, batch 200, pos+neg pairs)",
                   'events_per_step_per_gpu': G * REF_BATCH, 'reference_batch': REF_BATCH,
                   'sharding': 'whole reference batches round-robin over ranks; CSR + feature tables replicated',
                   'l2': 'flushed between timed steps (256 MiB write)', 'csr_build_s': round(build_s, 4)},
        'roofline': roofline,
        'e2e': {'value': events_total / (e2e_ms * 1e-3), 'unit': 'events/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': launches,
        'clocks': clk,
        'kernels': {k: {'ms': round(v[0], 3), 'launches': v[3], 'tflops': round(v[1] / max(v[0], 1e-9) / 1e9, 3),
                        'alg_gbs': round(v[2] / max(v[0], 1e-9) / 1e6, 1)} for k, v in per_kernel.items()},
        'score_checksum': checksum,
    }
    if rank == 0 and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        psd = {k: v.detach().cpu() for k, v in pred.state_dict().items()}
        ostep = oracle_dygformer(g, wl, sd, psd)
        nb = args.cpu_batches
        v = time_cpu(ostep, stream, nb)
        line['cpu_baseline'] = {'value': v, 'unit': 'events/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': f'{nb} reference batches of 200 events (pos+neg), oracle/ torch-CPU port, after 1 warm-up batch'}
        # sanity: the CPU port and the GPU path agree on the first timed batch
        hs = stream.rows(step_batches(W)[:1])
        want = torch.cat(ostep(*hs[:4]))
        got = dygformer_step(model, pred, *[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in hs[:4]]).cpu()
        line['parity_max_abs_err'] = float((want - got).abs().max())
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='dygformer_wiki', choices=['dygformer_wiki', 'dygformer_lastfm'])
    ap.add_argument('--batches-per-step', type=int, default=32)
    ap.add_argument('--cpu-batches', type=int, default=12)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, args.workload, wl)
    else:
        run_ours(args, args.workload, wl)


if __name__ == '__main__':
    main()
