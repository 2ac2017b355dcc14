#!/usr/bin/env python
"""Benchmark of the temporal neighbour-aggregation hot path (contract: DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Default workload = BASELINE.json configs[1]: DyGFormer link prediction on the synthetic Wikipedia-shaped
graph (157,474 events, 9,228 nodes, 172-d features), patch_size 2, max_input_sequence_length 64.
A step = one pass of the hot path over `--batches-per-step` reference batches of 200 events (each batch
keeps its own padding unit): negative draw (host, precomputed like the reference's seeded sampler) ->
first-hop search + pad -> co-occurrence -> patch projections -> 2 transformer layers -> link scores for
the positive and the negative pair of every event.

Other workloads (parity-test configs of BASELINE.json, reported in profiles/ and DESIGN.md, not the headline):
    tgat_myket, tgn_reddit, dygformer_lastfm, sampler_sweep
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
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def load_reference():
    """The UNMODIFIED reference modules of the path (scripts/install_reference.py copies them byte for byte into the git-ignored
    baseline/_ref/, which travels to the GPU box): {'utils': utils.utils, 'TGAT': models.TGAT, ...} or None when absent."""
    if not os.path.isfile(os.path.join(REF_DIR, 'models', 'MemoryModel.py')):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return {'utils': importlib.import_module('utils.utils'), 'modules': importlib.import_module('models.modules'),
            'TGAT': importlib.import_module('models.TGAT'), 'DyGFormer': importlib.import_module('models.DyGFormer'),
            'MemoryModel': importlib.import_module('models.MemoryModel')}


class ReferenceRunner:
    """One BASELINE config on the reference's own modules and sampler, driven the way evaluate_models_utils.py:67-138 drives
    them (model.eval(), torch.no_grad(), full-graph sampler, positive / negative pair per batch; memory models: negative call
    first).  ``device='cpu'`` is the CPU baseline, ``device='cuda:0'`` the "existing GPU path" (eager PyTorch on the same B200,
    SURVEY 2.1).  Weights: the state_dict of the GPU arm's model (same keys by construction)."""

    def __init__(self, ref, wl, sd, psd, device, sampler=None):
        g = wl.g
        t0 = time.perf_counter()
        self.sampler = sampler or ref['utils'].get_neighbor_sampler(g, 'recent', seed=1)
        self.sampler_build_s = time.perf_counter() - t0
        self.kind = type(wl).__name__
        if isinstance(wl, DyGFormerWL):
            m = ref['DyGFormer'].DyGFormer(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 50, wl.P, 2, 2, 0.1, wl.L, device)
        elif isinstance(wl, TGATWL):
            m = ref['TGAT'].TGAT(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 2, 2, 0.1, device)
        else:
            m = ref['MemoryModel'].MemoryModel(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 'TGN', 1, 2, 0.1, device=device)
        sd = {k: (torch.zeros_like(v) if ('node_memories' in k or 'node_last_updated_times' in k) else v) for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
        pred = ref['modules'].MergeLayer(172, 172, 172, 1)
        pred.load_state_dict(psd, strict=True)
        self.model = torch.nn.Sequential(m, pred).to(device).eval()
        self.device = device

    def step(self, src, dst, neg, t, eid):
        m, pred = self.model[0], self.model[1]
        with torch.no_grad():
            if self.kind == 'TGNWL':
                a, b = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=neg, node_interact_times=t, edge_ids=None,
                                                                  edges_are_positive=False, num_neighbors=10)
                c, d = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=dst, node_interact_times=t, edge_ids=eid,
                                                                  edges_are_positive=True, num_neighbors=10)
            elif self.kind == 'TGATWL':
                c, d = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=dst, node_interact_times=t, num_neighbors=20)
                a, b = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=neg, node_interact_times=t, num_neighbors=20)
            else:
                c, d = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=dst, node_interact_times=t)
                a, b = m.compute_src_dst_node_temporal_embeddings(src_node_ids=src, dst_node_ids=neg, node_interact_times=t)
            pos = pred(input_1=c, input_2=d).squeeze(dim=-1).sigmoid()
            negp = pred(input_1=a, input_2=b).squeeze(dim=-1).sigmoid()
            return torch.cat([pos, negp]).reshape(-1, 1)

    def copy_memory_state_to(self, other):
        """TGN: hand the warmed-up memory bank (memories, last-update times, pending messages) to another runner."""
        a, b = self.model[0].memory_bank, other.model[0].memory_bank
        b.node_memories.data.copy_(a.node_memories.data)
        b.node_last_updated_times.data.copy_(a.node_last_updated_times.data)
        b.node_raw_messages = type(a.node_raw_messages)(list)
        for v, lst in a.node_raw_messages.items():
            b.node_raw_messages[v] = [(mm[0].to(other.device), mm[1]) for mm in lst]


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p['bf16_tflops_sustained'], tensor_burst=p['bf16_tflops'], source='measured')
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source='fallback')


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML queried from a thread of this process every
    DYG_CLOCK_PERIOD_MS (default 100) ms (nvidia_ml_py).  A looping ``nvidia-smi -lms`` child was measured to stall the GPU for tens of milliseconds per
    sample (a 35 ms region of sampler launches read 65 ms whenever a sample fell inside it); it remains the fall-back when
    NVML cannot be loaded, started early so that only its periodic queries overlap the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.sm, self.mx, self.reasons, self.p, self.thread = [], [], set(), None, None
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # torch's device index counts within CUDA_VISIBLE_DEVICES; map it to the NVML index through the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            h = None
            for i in range(nv.nvmlDeviceGetCount()):
                hi = nv.nvmlDeviceGetHandleByIndex(i)
                u = nv.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            bits = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else 0x8,
                    'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4}
            self._stop = threading.Event()
            self.period = float(os.environ.get('DYG_CLOCK_PERIOD_MS', '100')) / 1e3

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        try:
                            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for nm, bit in bits.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    self._stop.wait(self.period)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = 'nvml'
            return
        except Exception:
            self.thread = None
        self.source = 'nvidia-smi'
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        t0 = time.time()
        while self.p is not None and time.time() - t0 < 5.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join()
            return {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                    'reasons': sorted(self.reasons), 'samples': len(self.sm), 'source': f'nvml, {self.period * 1e3:.0f} ms period'}
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
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi -lms 100'}


def dist_env():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def shard_batches(step, G, world, rank, nb):
    """Reference batches of one step for one rank: whole batches, round-robin over ranks (weak scaling)."""
    return [((step * G + j) * world + rank) % nb for j in range(G)]


class Stream:
    """Chronological reference batches with seeded random negatives (the reference's `random` negative mode,
    utils/utils.py:378-390, drawn once on the host before timing)."""

    def __init__(self, g, batch=None, region=0.15, seed=2, start=None):
        batch = batch or REF_BATCH
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
        return (g.src_node_ids[idx], g.dst_node_ids[idx], self.neg[nidx], g.node_interact_times[idx], g.edge_ids[idx])


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE.json config: GPU step through the package's public API + the oracle (CPU port) step."""
    name = ''
    dominant_bound = 'tensor'
    sequential = False        # TGN: batches form a dependency chain -> replicas only across GPUs
    default_G = 32
    cpu_batches = 12
    use_graph = False         # replay the step as a CUDA graph (fixed shapes, no host sync inside)

    def describe(self):
        raise NotImplementedError

    def build(self, dev):
        raise NotImplementedError

    def step(self, src, dst, neg, t, eid):
        raise NotImplementedError

    def oracle(self):
        raise NotImplementedError


def _predict(pred, a, b):
    from dyglib_b200 import ops
    return ops.mlp2([a, b], pred.fc1.weight.detach(), pred.fc1.bias.detach(), pred.fc2.weight.detach(), pred.fc2.bias.detach(),
                    act2=ops.ACT_SIGMOID)


class DyGFormerWL(Workload):
    def __init__(self, graph, P, L):
        self.name, self.P, self.L = graph, P, L
        self.cpu_batches = 12 if L <= 64 else 3
        self.default_G = 32 if L <= 64 else 8

    def describe(self):
        return f'{self.name} (DyGFormer P={self.P} L={self.L}, recent first-hop history, batch 200, pos+neg pairs)'

    def build(self, dev):
        from dyglib_b200.utils.utils import get_neighbor_sampler, set_random_seed
        from dyglib_b200.models.DyGFormer import DyGFormer
        from dyglib_b200.models.modules import MergeLayer
        self.g = g = make_config_graph(self.name)
        set_random_seed(0)
        t0 = time.perf_counter()
        self.sampler = get_neighbor_sampler(g, 'recent', device=dev)
        torch.cuda.synchronize()
        self.build_s = time.perf_counter() - t0
        self.model = DyGFormer(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 50, self.P, 2, 2, 0.1, self.L, dev).eval()
        self.pred = MergeLayer(172, 172, 172, 1).to(dev).eval()
        self.stream = Stream(g)

    def step(self, src, dst, neg, t, eid):
        # pos and neg pairs of every event; every reference batch is its own padding unit for pos and for neg
        s2, d2, t2 = torch.cat([src, src]), torch.cat([dst, neg]), torch.cat([t, t])
        es, ed = self.model.compute_src_dst_node_temporal_embeddings(s2, d2, t2, batch_size=REF_BATCH)
        return _predict(self.pred, es, ed)

    def oracle(self):
        from oracle.sampler import OracleSampler
        from oracle.models import OracleDyGFormer, merge_layer
        g = self.g
        sd = {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
        psd = {k: v.detach().cpu() for k, v in self.pred.state_dict().items()}
        samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
        m = OracleDyGFormer(sd, g.node_raw_features, g.edge_raw_features, samp, 50, self.P, 2, 2, self.L)

        def step(src, dst, neg, t, eid):
            out = []
            with torch.no_grad():
                for d in (dst, neg):
                    a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t)
                    out.append(torch.sigmoid(merge_layer(psd, '', a, b)))
            return torch.cat(out)
        return step


class TGATWL(Workload):
    name = 'tgat_myket'
    dominant_bound = 'hbm'
    default_G = 8
    cpu_batches = 2
    use_graph = True

    def describe(self):
        return 'tgat_myket (TGAT 2 layers, 20 recent neighbours, batch 200, pos+neg pairs; src embedding shared by both pairs)'

    def build(self, dev):
        from dyglib_b200.utils.utils import get_neighbor_sampler, set_random_seed
        from dyglib_b200.models.TGAT import TGAT
        from dyglib_b200.models.modules import MergeLayer
        self.g = g = make_config_graph(self.name)
        set_random_seed(0)
        t0 = time.perf_counter()
        self.sampler = get_neighbor_sampler(g, 'recent', device=dev)
        torch.cuda.synchronize()
        self.build_s = time.perf_counter() - t0
        self.model = TGAT(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 2, 2, 0.1, dev).eval()
        self.pred = MergeLayer(172, 172, 172, 1).to(dev).eval()
        self.stream = Stream(g)

    def step(self, src, dst, neg, t, eid):
        n = src.numel()
        emb = self.model.compute_node_temporal_embeddings(torch.cat([src, dst, neg]), torch.cat([t, t, t]), 2, 20)
        es = torch.cat([emb[:n], emb[:n]])
        return _predict(self.pred, es, emb[n:])

    def oracle(self):
        from oracle.sampler import OracleSampler
        from oracle.models import OracleTGAT, merge_layer
        g = self.g
        sd = {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
        psd = {k: v.detach().cpu() for k, v in self.pred.state_dict().items()}
        samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
        m = OracleTGAT(sd, g.node_raw_features, g.edge_raw_features, samp, 2, 2)

        def step(src, dst, neg, t, eid):
            out = []
            with torch.no_grad():
                for d in (dst, neg):       # the reference recomputes the src embedding for the negative pair
                    a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                    out.append(torch.sigmoid(merge_layer(psd, '', a, b)))
            return torch.cat(out)
        return step


class TGNWL(Workload):
    name = 'tgn_reddit'
    dominant_bound = 'hbm'
    sequential = True
    default_G = 1
    cpu_batches = 12
    use_graph = True

    def describe(self):
        return f'tgn_reddit (TGN 1 layer, 10 recent neighbours, last-message + GRU memory, batch {REF_BATCH} sequential, neg + pos roots in one pass)'

    def build(self, dev):
        from dyglib_b200.utils.utils import get_neighbor_sampler, set_random_seed
        from dyglib_b200.models.MemoryModel import MemoryModel
        from dyglib_b200.models.modules import MergeLayer
        self.g = g = make_config_graph(self.name)
        set_random_seed(0)
        t0 = time.perf_counter()
        self.sampler = get_neighbor_sampler(g, 'recent', device=dev)
        torch.cuda.synchronize()
        self.build_s = time.perf_counter() - t0
        self.model = MemoryModel(g.node_raw_features, g.edge_raw_features, self.sampler, 100, 'TGN', 1, 2, 0.1, device=dev).eval()
        self.model.memory_bank.__init_memory_bank__()
        self.pred = MergeLayer(172, 172, 172, 1).to(dev).eval()
        self.stream = Stream(g, start=0, region=1.0)

    def step(self, src, dst, neg, t, eid):
        # the reference loop's negative call then positive call (train_link_prediction.py:236-247) in one embedding pass
        # ... and the link predictor of train_link_prediction.py:243-244 inside the same launch (dyg_tgn_step)
        out = self.model.compute_pos_neg_temporal_embeddings(src, dst, neg, t, eid, 10, link_predictor=self.pred)
        pos, neg_ = out[4], out[5]
        base = pos._base
        if base is not None and base is neg_._base and base.numel() == pos.numel() + neg_.numel() and pos.data_ptr() == base.data_ptr():
            return base.reshape(-1, 1)          # the fused step wrote [pos | neg] into one buffer: no concatenation kernel
        return torch.cat([pos, neg_]).reshape(-1, 1)

    def oracle(self):
        from oracle.sampler import OracleSampler
        from oracle.models import OracleMemoryModel, merge_layer
        g = self.g
        sd = {k: v.detach().cpu() for k, v in self.model.state_dict().items()}
        for k in sd:
            if 'node_memories' in k or 'node_last_updated_times' in k:
                sd[k] = torch.zeros_like(sd[k])
        psd = {k: v.detach().cpu() for k, v in self.pred.state_dict().items()}
        samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
        m = OracleMemoryModel(sd, g.node_raw_features, g.edge_raw_features, samp, 'TGN', 1, 2)

        def step(src, dst, neg, t, eid):
            with torch.no_grad():
                a, b = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
                c, d = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
                return torch.sigmoid(merge_layer(psd, '', torch.cat([c, a]), torch.cat([d, b])))
        return step


def make_workload(name):
    if name == 'dygformer_wiki':
        return DyGFormerWL('dygformer_wiki', 2, 64)
    if name == 'dygformer_lastfm':
        return DyGFormerWL('dygformer_lastfm', 16, 512)
    if name == 'tgat_myket':
        return TGATWL()
    if name == 'tgn_reddit':
        return TGNWL()
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------ shared pieces
class Ctx:
    """Process-wide state of one bench run: rank / world, the device, torch.distributed (initialised once)."""

    def __init__(self):
        self.rank, self.world, self.local = dist_env()
        import torch.distributed as dist
        self.dist = dist
        if self.world > 1:
            torch.cuda.set_device(self.local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        else:
            torch.cuda.set_device(0)
        self.dev = torch.device('cuda', torch.cuda.current_device())
        self.pk = peaks()
        self.hold_graphs = []          # captured graphs with NCCL kernels are kept alive until the process exits
        self.need_hard_exit = False

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def finish(self):
        if self.world > 1:
            if self.need_hard_exit:
                # a captured graph holds NCCL kernels: tearing the communicator down under it hung at exit; leave together
                self.dist.barrier()
                torch.cuda.synchronize()
                sys.stdout.flush()
                sys.stderr.flush()
                os._exit(0)
            self.dist.destroy_process_group()


def workload_config(wl, G):
    """The `config` object of the JSON line: what is measured, identical for the GPU arm and for `--impl reference` (the CPU arm
    times a bounded sample of the same workload, stated in its cpu_baseline.sample)."""
    events_per_step = (1 if wl.sequential else G) * REF_BATCH
    return {'workload': wl.describe(), 'events_per_step_per_gpu': events_per_step, 'reference_batch': REF_BATCH,
            'sharding': ('replicas only (memory dependency chain)' if wl.sequential else
                         'whole reference batches round-robin over ranks; CSR + feature tables replicated'),
            'l2': 'GPU arm: flushed between timed steps (256 MiB write)'}


def cpu_arm(wl, sd, psd):
    """(step function, kind, what) of the CPU arm for workload ``wl``: the unmodified reference from baseline/_ref when it is
    there (kind "reference"), else the oracle port (kind "port")."""
    ref = load_reference()
    if ref is not None:
        r = ReferenceRunner(ref, wl, sd, psd, 'cpu')
        return r.step, 'reference', r
    return wl.oracle(), 'port', None


def time_cpu(step, stream, first, warm, nb):
    b = first
    for _ in range(warm):
        step(*stream.rows([b % stream.nb]))
        b += 1
    t0 = time.perf_counter()
    for _ in range(nb):
        step(*stream.rows([b % stream.nb]))
        b += 1
    return nb * REF_BATCH / (time.perf_counter() - t0)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (baseline/_ref, unmodified; the oracle port only when
    that copy is absent) with all host threads, on the GPU arm's config / metric / unit; a step is a bounded sample of the
    workload (a few reference batches) so that the run ends within minutes."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    wl = make_workload(args.workload)
    # weights: same constructors / seed as the GPU arm, built on CPU
    from dyglib_b200.utils.utils import set_random_seed
    set_random_seed(0)
    wl.g = g = make_config_graph(wl.name)
    from dyglib_b200.models.modules import MergeLayer
    if isinstance(wl, DyGFormerWL):
        from dyglib_b200.models.DyGFormer import DyGFormer
        wl.model = DyGFormer(g.node_raw_features, g.edge_raw_features, None, 100, 50, wl.P, 2, 2, 0.1, wl.L, 'cpu')
        wl.stream = Stream(g)
    elif isinstance(wl, TGATWL):
        from dyglib_b200.models.TGAT import TGAT
        wl.model = TGAT(g.node_raw_features, g.edge_raw_features, None, 100, 2, 2, 0.1, 'cpu')
        wl.stream = Stream(g)
    else:
        from dyglib_b200.models.MemoryModel import MemoryModel
        wl.model = MemoryModel(g.node_raw_features, g.edge_raw_features, None, 100, 'TGN', 1, 2, 0.1, device='cpu')
        wl.stream = Stream(g, start=0, region=1.0)
    wl.pred = MergeLayer(172, 172, 172, 1)
    sd = {k: v.detach().cpu() for k, v in wl.model.state_dict().items()}
    psd = {k: v.detach().cpu() for k, v in wl.pred.state_dict().items()}
    step, kind, _ = cpu_arm(wl, sd, psd)
    per_step = 1 if isinstance(wl, (TGATWL,)) or (isinstance(wl, DyGFormerWL) and wl.L > 64) else 2     # bounded sample per step
    G = args.batches_per_step or wl.default_G
    b = 0
    for _ in range(args.warmup + (40 if isinstance(wl, TGNWL) else 0)):   # memory model: realistic pending set
        step(*wl.stream.rows([b % wl.stream.nb]))
        b += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            step(*wl.stream.rows([b % wl.stream.nb]))
            b += 1
    total = time.perf_counter() - t0
    value = args.steps * per_step * REF_BATCH / total
    cores = torch.get_num_threads()
    what = ('unmodified reference modules + NeighborSampler from baseline/_ref (python sampler loop single-core, torch ops '
            f'{cores} threads)') if kind == 'reference' else f'oracle/ = torch-CPU port of the reference path (torch ops {cores} threads)'
    print(json.dumps({
        'impl': 'reference', 'metric': 'link-pred events/sec', 'value': value, 'unit': 'events/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(wl, G),
        'cpu_baseline': {'value': value, 'unit': 'events/s', 'cores': cores, 'kind': kind,
                         'sample': f'{per_step} reference batch(es) of {REF_BATCH} events (pos+neg) per step; {what}'},
        'e2e': {'value': value, 'unit': 'events/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def kernel_traffic(workload, kernel, G):
    """DRAM bytes per launch (GB) of the dominant kernel from the committed ncu --set full capture of the same workload at the
    same launch sizes; None when there is no capture for this configuration."""
    for name in ('r02_kernel_traffic.json', 'r01_kernel_traffic.json'):
        try:
            d = json.load(open(os.path.join(ROOT, 'profiles', name)))[workload]
            if d.get('batches_per_step') == G and d.get(kernel) is not None:
                return d.get(kernel), name
        except Exception:
            pass
    return None, None


# ------------------------------------------------------------------------------------------------ our arm
def measure_model(ctx, args, name, K, W, with_cpu=True):
    """One BASELINE model config on this rank's GPU: device-resident `value`, end-to-end `e2e`, per-kernel roofline pass, and on
    rank 0 of a 1-GPU run the parity check, the CPU baseline and the eager-PyTorch-on-B200 baseline.  Returns the JSON line."""
    rank, world, dev, dist, pk = ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.pk
    from dyglib_b200 import ops
    wl = make_workload(name)
    wl.build(dev)
    stream = wl.stream
    G = args.batches_per_step or wl.default_G
    if wl.sequential:
        # dependency chain through the memory: every rank runs an independent replica of the same stream
        batches_of = lambda i: [i % stream.nb]   # noqa: E731
    else:
        batches_of = lambda i: shard_batches(i, G, world, rank, stream.nb)   # noqa: E731

    def reset():
        if wl.sequential:
            wl.model.memory_bank.__init_memory_bank__()

    total_steps = 3 * (W + K) if wl.sequential else (W + K)
    host_steps = [stream.rows(batches_of(i)) for i in range(total_steps)]
    to_dev = lambda hs: tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in hs)   # noqa: E731
    dev_steps = [to_dev(hs) for hs in host_steps]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(run_step, first):
        evs = []
        for i in range(first, first + K):
            flush.zero_()                                  # L2 flush between timed iterations (outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run_step(i)
            e1.record()
            evs.append((e0, e1))
        ctx.barrier()
        return sum(a.elapsed_time(b) for a, b in evs), out

    step_fn = wl.step
    graphed = None
    if wl.use_graph and not args.no_graph:
        from dyglib_b200.utils.graph import GraphedStep
        graphed = GraphedStep(wl.step, dev_steps[0], warmup=2, after_warmup=reset)
        step_fn = graphed
    with torch.no_grad():
        # ---------------- device-resident inputs
        reset()
        clocks = ClockSampler(torch.cuda.current_device())   # started before the warm-up: see ClockSampler.__init__
        for i in range(W):
            step_fn(*dev_steps[i])
        ctx.barrier()
        launches0 = ops.launch_count
        total_ms, scores = timed(lambda i: step_fn(*dev_steps[i]), W)
        scores = scores.clone()
        launches = ops.launch_count - launches0
        clk = clocks.stop()
        # ---------------- end to end: pinned host buffers in, host scores out, copies inside the timed region
        pinned = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in hs) for hs in host_steps]
        out_host = torch.empty(tuple(scores.shape), dtype=torch.float32).pin_memory()
        base = (W + K) if wl.sequential else 0

        def e2e_step(i):
            sc = graphed(*pinned[i]) if graphed is not None else wl.step(*[a.to(dev, non_blocking=True) for a in pinned[i]])
            out_host.copy_(sc, non_blocking=True)
            return sc
        for i in range(base, base + W):
            e2e_step(i)
        ctx.barrier()
        e2e_ms, _ = timed(e2e_step, base + W)
        h2d = sum(a.numel() * a.element_size() for a in pinned[0])
        d2h = out_host.numel() * 4
        # ---------------- roofline pass: same steps again, every launch bracketed by CUDA events on its stream
        base = 2 * (W + K) if wl.sequential else 0
        for i in range(base, base + W):
            wl.step(*dev_steps[i])
        ops.PROFILE = []
        for i in range(base + W, base + W + K):
            flush.zero_()
            wl.step(*dev_steps[i])
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    if world > 1:
        tt = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = tt.tolist()
        # final score gather over NCCL (outside the timed region)
        gathered = torch.empty((world,) + tuple(scores.shape), device=dev, dtype=scores.dtype)
        dist.all_gather_into_tensor(gathered, scores.contiguous())
        checksum = float(gathered.double().sum().item())
    else:
        checksum = float(scores.double().sum().item())
    per_kernel = {}
    for kname, e0, e1, fl, by in prof:
        d = per_kernel.setdefault(kname, [0.0, 0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1)
        d[1] += fl
        d[2] += by
        d[3] += 1
    events_per_step = (1 if wl.sequential else G) * REF_BATCH
    events_total = K * events_per_step * world
    value = events_total / (total_ms * 1e-3)
    kname, (ms, fl, by, cnt) = max(per_kernel.items(), key=lambda kv: kv[1][0])
    share = ms / max(sum(v[0] for v in per_kernel.values()), 1e-9)
    if kname in TENSOR_KERNELS:
        achieved = fl / (ms * 1e-3) / 1e12
        roofline = {'kernel': kname, 'bound': 'tensor', 'achieved': achieved, 'peak': pk['tensor'], 'unit': 'TFLOP/s',
                    'frac': achieved / pk['tensor'], 'traffic': None, 'peak_source': pk['source'] + ' (bf16 sustained)',
                    'note': 'algorithmic fp32 flops (sum of 2MNK) vs the measured bf16 peak; products run as BF16x3 (3 bf16 MMAs '
                            'each, fp32 parity), so the fp32-equivalent ceiling is peak / 3',
                    'mma_frac': 3.0 * achieved / pk['tensor'] if kname not in FFMA_KERNELS else None}
    else:
        achieved = by / (ms * 1e-3) / 1e9
        roofline = {'kernel': kname, 'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm'], 'unit': 'GB/s',
                    'frac': achieved / pk['hbm'], 'traffic': None, 'peak_source': pk['source'],
                    'note': 'algorithmic bytes = gathered rows + indices + query/result vectors per launch'}
    roofline.update({'launches': cnt, 'avg_launch_us': 1e3 * ms / cnt, 'share_of_timed_kernels': share})
    roofline['traffic'], src_file = kernel_traffic(wl.name, kname, G)
    if roofline['traffic'] is not None:
        roofline['traffic_unit'] = f'GB per launch (dram read + write, committed ncu capture: profiles/{src_file})'
    line = {
        'metric': 'link-pred events/sec', 'value': value, 'unit': 'events/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(wl, G),
        'details': {'csr_build_s': round(wl.build_s, 4),
                    'launch': 'CUDA graph replay of the captured step' if graphed is not None else 'direct launches',
                    'e2e_inputs': 'pinned host numpy buffers (ids, times, precomputed seeded negatives) -> H2D inside the timed region'},
        'roofline': roofline,
        'e2e': {'value': events_total / (e2e_ms * 1e-3), 'unit': 'events/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': launches, 'clocks': clk,
        'kernels': {k: {'ms': round(v[0], 3), 'launches': v[3], 'tflops': round(v[1] / max(v[0], 1e-9) / 1e9, 3),
                        'alg_gbs': round(v[2] / max(v[0], 1e-9) / 1e6, 1)} for k, v in per_kernel.items()},
        'score_checksum': checksum,
    }
    line['step_roofline'] = step_roofline(wl, pk, events_per_step, total_ms / K, launches // max(K, 1))
    if rank == 0 and world == 1 and with_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu() for k, v in wl.model.state_dict().items()}
        psd = {k: v.detach().cpu() for k, v in wl.pred.state_dict().items()}
        cstep, kind, runner = cpu_arm(wl, sd, psd)
        nb = args.cpu_batches or max(1, wl.cpu_batches * 200 // REF_BATCH)
        first = 0 if wl.sequential else batches_of(W)[0]
        # memory models: the reference recomputes every pending node per call, so its cost depends on how many nodes hold a
        # pending message (>= 40 warm-up batches), and the memory path (last-message choice, GRU) is only exercised once
        # there is state: parity is taken AFTER the warm-up batches, on three consecutive batches
        cpu_warm = max(0, 40 * 200 // REF_BATCH) if wl.sequential else 0
        reset()
        with torch.no_grad():
            errs = []
            for b in range(cpu_warm + 3):
                hs = stream.rows([(first + b) % stream.nb])
                want = cstep(*hs) if (wl.sequential or b >= cpu_warm) else None
                got = wl.step(*to_dev(hs)).cpu() if (wl.sequential or b >= cpu_warm) else None
                if b >= cpu_warm:
                    errs.append(float((want.cpu() - got).abs().max()))
        line['parity_max_abs_err'] = max(errs)
        line['parity'] = {'against': 'unmodified reference (baseline/_ref) on CPU' if kind == 'reference' else 'oracle/ port on CPU',
                          'batches': 3, 'after_batches_of_state': cpu_warm, 'max_abs_err_link_probability': max(errs)}
        v = time_cpu(cstep, stream, first + cpu_warm + 3, 0 if wl.sequential else 1, nb)
        line['cpu_baseline'] = {'value': v, 'unit': 'events/s', 'cores': torch.get_num_threads(), 'kind': kind,
                                'sample': f'{nb} reference batches of {REF_BATCH} events (pos+neg) after {cpu_warm + 3 if wl.sequential else 1} '
                                          f'warm-up batch(es); ' + ('unmodified reference modules + python NeighborSampler (baseline/_ref)'
                                                                    if kind == 'reference' else 'oracle/ torch-CPU port')}
        # the "existing GPU path": the same unmodified reference modules run eagerly by PyTorch on this B200
        ref = load_reference()
        if ref is not None and not args.no_eager:
            try:
                gr = ReferenceRunner(ref, wl, sd, psd, f'cuda:{torch.cuda.current_device()}', sampler=runner.sampler)
                if wl.sequential:
                    runner.copy_memory_state_to(gr)
                ne = max(2, min(nb, 6))
                f0 = first + cpu_warm + 3 + nb + 1
                gr.step(*stream.rows([f0 % stream.nb]))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for b in range(1, ne + 1):
                    sc = gr.step(*stream.rows([(f0 + b) % stream.nb]))
                    sc.cpu()
                ev = ne * REF_BATCH / (time.perf_counter() - t0)
                line['gpu_eager_baseline'] = {'value': ev, 'unit': 'events/s', 'kind': 'reference modules, device=cuda (eager PyTorch '
                                              f'{torch.__version__}); its sampler / padding / co-occurrence loops run on the host as in the reference',
                                              'sample': f'{ne} reference batches after 1 warm-up batch, wall clock incl. the score read-back'}
                del gr
            except Exception as e:   # noqa: BLE001
                line['gpu_eager_baseline'] = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
        else:
            line['gpu_eager_baseline'] = {'unavailable': 'baseline/_ref absent' if ref is None else 'disabled (--no-eager)'}
    return line


TENSOR_KERNELS = ('linear_kernel', 'linear_tc_kernel', 'gemm_bf16x3_kernel', 'ln_ffn_bf16x3_kernel', 'qkv_ln_gemm_kernel',
                  'seq_attention_tc5_kernel', 'tgn_step_kernel', 'gru_update_kernel')
FFMA_KERNELS = ('linear_kernel', 'gru_update_kernel')


def step_roofline(wl, pk, events_per_step, ms_per_step, launches_per_step):
    """Step-level roofline (SURVEY 8d per-event figures): the time the whole step would take at the HBM / tensor roofline of its
    algorithmic bytes / flops, over the measured step time."""
    per_event = {   # (bytes, flops) per event; TGAT counted with the 3 roots per event the step embeds
        'tgat_myket': (3 * 616e3, 3 * 22 * 10.19e6), 'dygformer_wiki': (356e3, 274e6), 'dygformer_lastfm': (2.85e6, 380e6),
        'tgn_reddit': (4 * (10 * 3 * 172 * 4 + 2 * 172 * 4 + 288) + 2 * (2 * 616 * 4 + 4 * 172 * 4), 23e6)}[wl.name]
    hbm_us = events_per_step * per_event[0] / (pk['hbm'] * 1e9) * 1e6
    tensor_us = events_per_step * per_event[1] / (pk['tensor'] * 1e12) * 1e6
    bound = 'tensor' if tensor_us > hbm_us else 'hbm'
    out = {'bound': bound if not wl.sequential else 'latency', 'alg_bytes_per_event': per_event[0], 'alg_flops_per_event': per_event[1],
           'hbm_time_us': hbm_us, 'tensor_time_us': tensor_us, 'step_us': 1e3 * ms_per_step,
           'frac': max(hbm_us, tensor_us) / (1e3 * ms_per_step), 'launches_per_step': launches_per_step}
    if wl.sequential:
        out['frac'] = hbm_us / (1e3 * ms_per_step)
        out['note'] = 'B=200 dependency chain: the step time is launches x per-phase latency, not bytes / bandwidth (SURVEY 7.3(6))'
    return out


def compact(line):
    """The per-workload record kept inside the default line's `workloads` object."""
    keep = ('metric', 'value', 'unit', 'ms_per_step', 'steps', 'e2e', 'gpu_launches', 'parity_max_abs_err', 'cpu_baseline',
            'gpu_eager_baseline', 'step_roofline', 'strategies', 'allreduce_ms', 'ranks_in_sync', 'parity_recent_bit_exact', 'error')
    out = {k: line[k] for k in keep if k in line}
    if 'roofline' in line:
        out['roofline'] = {k: v for k, v in line['roofline'].items() if k not in ('note', 'traffic_unit', 'peak_source')}
    if 'config' in line:
        out['workload'] = line['config'].get('workload')
    for k in ('cpu_baseline', 'gpu_eager_baseline'):
        if k in out and isinstance(out[k], dict):
            out[k] = {kk: vv for kk, vv in out[k].items() if kk not in ('sample',) or len(str(vv)) < 90}
    if 'step_roofline' in out:
        out['step_roofline'] = {k: v for k, v in out['step_roofline'].items() if k != 'note'}
    return out


def run_ours(ctx, args):
    import gc
    K, W = args.steps, args.warmup
    line = measure_model(ctx, args, args.workload, K, W)
    if args.workload == 'dygformer_wiki' and not args.only_headline and not args.batches_per_step:
        # the default command carries every BASELINE config (headline stays configs[1]); each with its own value / e2e /
        # roofline / parity / cpu_baseline, on fewer steps so that the whole run stays within minutes
        line['workloads'] = {}
        Kx = max(3, min(K, 10))
        for name in ('tgat_myket', 'tgn_reddit', 'dygformer_lastfm', 'sampler_sweep') + (('tgat_train',) if ctx.world > 1 else ()):
            gc.collect()
            torch.cuda.empty_cache()
            try:
                if name == 'sampler_sweep':
                    sub = measure_sampler_sweep(ctx, args, 5, 3)
                elif name == 'tgat_train':
                    sub = measure_train(ctx, args, max(K, 10), W)
                else:
                    sub = measure_model(ctx, args, name, Kx, W)
            except Exception as e:   # noqa: BLE001   (an extra workload must not take the headline down)
                import traceback
                traceback.print_exc(file=sys.stderr)
                sub = {'error': f'{type(e).__name__}: {e}'[:300]}
            line['workloads'][name] = compact(sub)
            if args.save_dir and ctx.rank == 0:
                os.makedirs(args.save_dir, exist_ok=True)
                json.dump(sub, open(os.path.join(args.save_dir, f'{name}_{ctx.world}gpu.json'), 'w'))
    return line


# ------------------------------------------------------------------------------------------------ training configuration
def measure_train(ctx, args, K, W):
    """TGAT training step (train_link_prediction.py:230-257: pos + neg embeddings, BCE, backward, Adam) on one reference batch
    of 200 events per rank per step; data parallel over ranks with ONE gradient all-reduce (NCCL) per step over the flat
    gradient bucket (the only collective on the path).  Dropout 0.1 as in the reference's defaults."""
    rank, world, dev, dist, pk = ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.pk
    from dyglib_b200 import ops
    from dyglib_b200.utils.dist import GradBucket
    wl = TGATWL()
    wl.build(dev)
    model, pred, stream = wl.model.train(), wl.pred.train(), wl.stream
    params = list(model.parameters()) + list(pred.parameters())
    if world > 1:                                   # same initial weights on every rank
        for p_ in params:
            dist.broadcast(p_.data, 0)
    bucket = GradBucket(params)
    opt = torch.optim.Adam(bucket.params, lr=1e-4, capturable=not args.no_graph)
    host_steps = [stream.rows(shard_batches(i, 1, world, rank, stream.nb)) for i in range(W + K)]
    to_dev = lambda hs: tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in hs)   # noqa: E731
    dev_steps = [to_dev(hs) for hs in host_steps]
    pinned = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in hs) for hs in host_steps]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def train_step(src, dst, neg, t, eid):
        bucket.zero()
        # the four root sets of the reference's two calls (models/TGAT.py:48-64 for (src, dst) and for (src, neg); the src
        # embedding of the negative pair is recomputed under its own dropout masks) go through ONE recursion of 4 B roots
        B = src.numel()
        emb = model.compute_node_temporal_embeddings(torch.cat([src, dst, src, neg]), torch.cat([t, t, t, t]), 2, 20)
        pos = pred(emb[:B], emb[B:2 * B]).squeeze(dim=-1).sigmoid()
        negp = pred(emb[2 * B:3 * B], emb[3 * B:]).squeeze(dim=-1).sigmoid()
        predicts = torch.cat([pos, negp], dim=0)
        labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
        loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
        loss.backward()
        bucket.allreduce()
        opt.step()
        return loss.detach()

    barrier = ctx.barrier

    def timed(run_step):
        evs = []
        for i in range(W, W + K):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = run_step(i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs), out

    step_fn = train_step
    if not args.no_graph:
        # the whole step (sampling, forward, backward, all-reduce, Adam) replayed as one CUDA graph: at B=200 the direct
        # step is bound by the host issuing ~130 kernels of ours plus the autograd / optimizer ops
        from dyglib_b200.utils.graph import GraphedStep
        step_fn = GraphedStep(train_step, dev_steps[0], warmup=3, grad=True)
        ctx.hold_graphs.append(step_fn)
        ctx.need_hard_exit = ctx.need_hard_exit or world > 1
    clocks = ClockSampler(torch.cuda.current_device())
    for i in range(W):
        step_fn(*dev_steps[i])
    bucket.check_views()
    barrier()
    launches0 = ops.launch_count
    total_ms, loss = timed(lambda i: step_fn(*dev_steps[i]))
    launches = ops.launch_count - launches0
    clk = clocks.stop()

    def e2e_step(i):
        ls = step_fn(*pinned[i]) if step_fn is not train_step else train_step(*[a.to(dev, non_blocking=True) for a in pinned[i]])
        loss_host.copy_(ls.reshape(1), non_blocking=True)
        return ls
    for i in range(W):
        e2e_step(i)
    barrier()
    e2e_ms, _ = timed(e2e_step)
    h2d = sum(a.numel() * a.element_size() for a in pinned[0])
    # per-kernel pass (our launches only; the dense backward GEMMs are library calls and are not in this list)
    ops.PROFILE = []
    for i in range(W, W + K):
        flush.zero_()
        train_step(*dev_steps[i])
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    # gradient all-reduce alone (device time of the collective on the flat bucket)
    ar_ms = None
    if world > 1:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dist.all_reduce(bucket.flat)
        e1.record()
        torch.cuda.synchronize()
        ar_ms = e0.elapsed_time(e1) / 10
        tt = torch.tensor([total_ms, e2e_ms, ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, ar_ms = tt.tolist()
        # the ranks must have stayed in lock step: identical weights after the same averaged gradients
        flatw = torch.cat([p_.detach().reshape(-1) for p_ in bucket.params])
        ref = flatw.clone()
        dist.broadcast(ref, 0)
        in_sync = bool((ref == flatw).all().item())
    else:
        in_sync = True
    per_kernel = {}
    for name, e0, e1, fl, by in prof:
        d = per_kernel.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1)
        d[1] += fl
        d[2] += by
        d[3] += 1
    name, (ms, fl, by, cnt) = max(per_kernel.items(), key=lambda kv: kv[1][0])
    if by > 0 and name.startswith('temporal_attend'):
        achieved = by / (ms * 1e-3) / 1e9
        roofline = {'kernel': name, 'bound': 'hbm', 'achieved': achieved, 'peak': pk['hbm'], 'unit': 'GB/s', 'frac': achieved / pk['hbm'],
                    'traffic': None, 'peak_source': pk['source']}
    else:
        achieved = fl / (ms * 1e-3) / 1e12
        roofline = {'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': pk['tensor'], 'unit': 'TFLOP/s',
                    'frac': achieved / pk['tensor'], 'traffic': None, 'peak_source': pk['source'] + ' (bf16 sustained)'}
    roofline.update({'launches': cnt, 'avg_launch_us': 1e3 * ms / cnt,
                     'share_of_timed_kernels': ms / max(sum(v[0] for v in per_kernel.values()), 1e-9),
                     'note': 'B=200 training step: launch / latency bound (about 100 launches of small kernels per step)'})
    events_total = K * REF_BATCH * world
    line = {
        'metric': 'link-pred training events/sec', 'value': events_total / (total_ms * 1e-3), 'unit': 'events/s', 'n_gpus': world,
        'steps': K, 'warmup': W, 'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'tgat_train (TGAT 2 layers, 20 recent neighbours, dropout 0.1, batch 200 per rank, pos+neg, BCE, Adam lr 1e-4)',
                   'events_per_step_per_gpu': REF_BATCH, 'reference_batch': REF_BATCH,
                   'sharding': 'data parallel: one reference batch per rank per step, one NCCL all-reduce of the flat fp32 gradient '
                               'bucket per step; CSR + feature tables replicated',
                   'grad_elements': int(bucket.flat.numel()), 'l2': 'flushed between timed steps (256 MiB write)',
                   'csr_build_s': round(wl.build_s, 4),
                   'launch': 'direct launches' if step_fn is train_step else 'CUDA graph replay of the captured training step'},
        'roofline': roofline,
        'e2e': {'value': events_total / (e2e_ms * 1e-3), 'unit': 'events/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
        'gpu_launches': launches, 'clocks': clk, 'allreduce_ms': ar_ms, 'ranks_in_sync': in_sync, 'final_loss': float(loss.item()),
        'kernels': {k: {'ms': round(v[0], 3), 'launches': v[3]} for k, v in per_kernel.items()},
    }
    if rank == 0 and world == 1 and args.workload == 'tgat_train':
        # CPU arm: the oracle port's training step (autograd through the torch-CPU restatement) on the same batches
        torch.set_num_threads(os.cpu_count() or 1)
        from oracle.sampler import OracleSampler
        from oracle.models import OracleTGAT, merge_layer
        g = wl.g
        sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
        psd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in pred.state_dict().items()}
        samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
        om = OracleTGAT(sd, g.node_raw_features, g.edge_raw_features, samp, 2, 2)
        nb = args.cpu_batches or 2
        t0 = None
        for b in range(nb + 1):
            if b == 1:
                t0 = time.perf_counter()
            src, dst, neg, t, _ = stream.rows([b % stream.nb])
            ps, pd = om.compute_src_dst_node_temporal_embeddings(src, dst, t, 20)
            ns, nd = om.compute_src_dst_node_temporal_embeddings(src, neg, t, 20)
            pr = torch.cat([merge_layer(psd, '', ps, pd).squeeze(-1).sigmoid(), merge_layer(psd, '', ns, nd).squeeze(-1).sigmoid()])
            lb = torch.cat([torch.ones(len(src)), torch.zeros(len(src))])
            torch.nn.functional.binary_cross_entropy(pr, lb).backward()
        v = nb * REF_BATCH / (time.perf_counter() - t0)
        line['cpu_baseline'] = {'value': v, 'unit': 'events/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': f'{nb} training batches of 200 events (forward + backward, no dropout), oracle/ torch-CPU port, '
                                          'after 1 warm-up batch'}
    return line


# ------------------------------------------------------------------------------------------------ sampler sweep
def device_power_law_graph(E, nu, ni, seed, dev):
    """Config 5: 1e8-event power-law bipartite stream, strictly increasing integer times;
    generated on the device (the host generator would take minutes at this size)."""
    gen = torch.Generator(device=dev).manual_seed(seed)

    def zipf(n_items, alpha, size):
        cdf = torch.cumsum(torch.arange(1, n_items + 1, device=dev, dtype=torch.float64) ** (-alpha), 0)
        cdf /= cdf[-1].clone()
        out = torch.empty(size, dtype=torch.int64, device=dev)
        for s in range(0, size, 1 << 25):
            u = torch.rand(min(1 << 25, size - s), generator=gen, device=dev, dtype=torch.float64)
            out[s:s + u.numel()] = torch.searchsorted(cdf, u, right=True).clamp_(max=n_items - 1)
        return out
    pu = torch.randperm(nu, generator=gen, device=dev)
    pi = torch.randperm(ni, generator=gen, device=dev)
    src = 1 + pu[zipf(nu, 0.8, E)]
    dst = 1 + nu + pi[zipf(ni, 1.0, E)]
    t = torch.cumsum(torch.randint(1, 4, (E,), generator=gen, device=dev), 0).double()
    eid = torch.arange(1, E + 1, device=dev)
    return src, dst, eid, t, nu + ni + 1


def sweep_traffic(E, Q, k):
    """DRAM bytes per sample_recent launch (GB) from the committed ncu --set full capture of the same sweep configuration."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_sampler_traffic.json')))
        if (d['events'], d['queries'], d['k']) == (E, Q, k):
            return d['sample_recent_kernel']['dram_read_gb'] + d['sample_recent_kernel']['dram_write_gb']
    except Exception:
        pass
    return None


def measure_sampler_sweep(ctx, args, K, W):
    rank, world, dev, dist, pk = ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.pk
    from dyglib_b200 import ops
    from dyglib_b200.utils.utils import NeighborSampler
    E, Q, k = args.events, args.queries, 20
    nu, ni = max(8, int(E * 0.08)), max(4, int(E * 0.02))
    src, dst, eid, t, num_nodes = device_power_law_graph(E, nu, ni, 5, dev)
    t0 = time.perf_counter()
    samplers = {}
    for strat in ('recent', 'uniform', 'time_interval_aware'):
        if strat == 'recent':
            samplers[strat] = NeighborSampler(None, strat, 1e-6, 0, dev, 'philox', 'device', _edges=(src, dst, eid, t, num_nodes, True))
            torch.cuda.synchronize()
            build_s = time.perf_counter() - t0
        else:   # share the CSR, add the tia tables once
            s = object.__new__(NeighborSampler)
            s.__dict__.update(samplers['recent'].__dict__)
            s.sample_neighbor_strategy = strat
            if strat == 'time_interval_aware':
                s.time_scaling_factor = 1e-6
                s._build_tia('device')
            samplers[strat] = s
    # queries: event uniform in the last 30 %, endpoint by coin flip, time = event time; contiguous shard per rank
    gen = torch.Generator(device=dev).manual_seed(1234)
    qe = torch.randint(int(E * 0.7), E, (Q * world,), generator=gen, device=dev)
    coin = torch.rand(Q * world, generator=gen, device=dev) < 0.5
    nodes_all = torch.where(coin, src[qe], dst[qe])
    times_all = t[qe]
    nodes = nodes_all[rank * Q:(rank + 1) * Q].contiguous()
    times = times_all[rank * Q:(rank + 1) * Q].contiguous()
    del src, dst, eid, qe, coin
    base = samplers['recent']
    cnt = base.count_before_device(nodes, times).double()
    deg = (base.indptr[nodes + 1] - base.indptr[nodes]).double()
    m = torch.clamp(cnt, max=k)
    log_deg = torch.ceil(torch.log2(deg + 1))
    log_cnt = torch.ceil(torch.log2(cnt + 1))
    alg = {'recent': float((32 + 8 * log_deg + 16 * m + 20 * k).sum()),
           'uniform': float((32 + 8 * log_deg + 16 * torch.clamp(cnt, max=1) * k + 20 * k).sum()),
           'time_interval_aware': float((32 + 8 * log_deg + 16 * torch.clamp(cnt, max=1) * k + 8 * k * log_cnt + 20 * k).sum())}
    res = {}
    W = max(W, 3)
    for strat, s in samplers.items():
        clocks = ClockSampler(torch.cuda.current_device()) if strat == 'recent' else None
        for _ in range(W):
            # keep the previous outputs alive while the next ones are allocated, exactly as the timed loop does: otherwise the
            # second set of (n, k) output buffers (6.7 GB) is cudaMalloc'ed inside the first timed region
            out = s.get_historical_neighbors_device(nodes, times, k)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record()
        for i in range(K):
            out = s.get_historical_neighbors_device(nodes, times, k)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[K])
        per_launch = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
        if clocks is not None:
            clk = clocks.stop()
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        gbs = alg[strat] * K / (ms * 1e-3) / 1e9
        res[strat] = {'queries_per_s': Q * world * K / (ms * 1e-3), 'ms_per_launch': ms / K, 'ms_per_launch_median': per_launch[K // 2], 'alg_bytes_per_query': alg[strat] / Q,
                      'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / pk['hbm'],
                      'checksum': int(out[1].sum().item())}
    line = {'metric': 'sampler queries/s', 'value': res['recent']['queries_per_s'], 'unit': 'queries/s', 'n_gpus': world,
            'steps': K, 'warmup': W, 'ms_per_step': res['recent']['ms_per_launch'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64 compare / int32 ids', 'data': 'synthetic',
            'config': {'workload': f'sampler_sweep: {E} events, {num_nodes} nodes power-law, {Q} queries per GPU in the last 30 %, k={k}',
                       'csr_build_s': round(build_s, 3), 'csr_bytes': int(base.halfedges.numel() * 8 + base.indptr.numel() * 8),
                       'fence_bytes': int(base.fence.numel() * 8) if base.fence is not None else 0,
                       'l2': 'CSR (>= 3 GB at 1e8 events) and outputs far exceed L2', 'random_strategies': 'philox (throughput mode, non-parity)'},
            'roofline': {'kernel': 'sample_recent_kernel', 'bound': 'hbm', 'achieved': res['recent']['achieved_gbs'], 'peak': pk['hbm'],
                         'unit': 'GB/s', 'frac': res['recent']['frac_of_hbm_peak'], 'traffic': sweep_traffic(E, Q, k), 'peak_source': pk['source'],
                         'note': 'achieved = algorithmic bytes (SURVEY 8d formula, summed over the actual queries) / launch time; traffic = '
                                 'dram read + write bytes per launch from the committed ncu capture (profiles/r01_sampler_ncu.md), GB'},
            'strategies': res, 'gpu_launches': K * 3, 'clocks': clk}
    # ---- bit-exact replay mode of the random strategies (counts D2H -> the reference's RandomState stream on the host -> gather on
    # the device): a single sequential stream by construction, timed by wall clock on a bounded number of queries
    replay_samplers = {}
    for strat, nq_r in (('uniform', min(Q, 1 << 20)), ('time_interval_aware', min(Q, 256))):
        r = object.__new__(NeighborSampler)
        r.__dict__.update(samplers[strat].__dict__)
        r.rng, r.seed = 'numpy_replay', 0
        r._prob_host = None
        r.reset_random_state()
        replay_samplers[strat] = r
        if strat == 'time_interval_aware':
            # its host softmax needs the probability table of the queried rows only: fetch the needed slices lazily
            r._prob_host = _LazyDeviceVector(samplers[strat].tia_prob)
        try:
            r.get_historical_neighbors_device(nodes[:min(nq_r, 64)], times[:min(nq_r, 64)], k)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            r.get_historical_neighbors_device(nodes[:nq_r], times[:nq_r], k)
            torch.cuda.synchronize()
            res[strat]['replay'] = {'queries': nq_r, 'queries_per_s': nq_r / (time.perf_counter() - t1),
                                    'note': 'bit-exact mode: host MT19937 stream (one sequential stream), wall clock'}
        except ValueError as e:
            res[strat]['replay'] = {'queries': nq_r, 'raises': str(e)[:80],
                                    'note': 'RandomState.choice(p=float32 softmax) rejects hub rows, exactly as utils/utils.py:183-187 does'}
    if rank == 0 and world == 1 and args.cpu_queries > 0:
        # CPU baseline: the UNMODIFIED query code of the reference's NeighborSampler (utils/utils.py:130-214) over per-node slice views
        # of the CSR arrays (its constructor cannot build 1e8 events: BASELINE.md 3.5); the oracle port when baseline/_ref is absent
        rec = base.halfedges[:base.num_half_edges].cpu().numpy()
        t_all = np.ascontiguousarray(rec[:, 0])
        ints = np.ascontiguousarray(rec[:, 1]).view(np.int32).reshape(-1, 2)
        nbr_all, eid_all = ints[:, 0].astype(np.int64), ints[:, 1].astype(np.int64)
        indptr_h = base.indptr.cpu().numpy()
        ref = load_reference()
        if ref is not None:
            o = object.__new__(ref['utils'].NeighborSampler)
            o.nodes_neighbor_ids, o.nodes_edge_ids = _NodeSlices(nbr_all, indptr_h), _NodeSlices(eid_all, indptr_h)
            o.nodes_neighbor_times = _NodeSlices(t_all, indptr_h)
            kind = 'reference'
        else:
            from oracle.sampler import OracleSampler
            o = object.__new__(OracleSampler)
            o.t, o.nbr, o.eid, o.indptr, o.num_nodes = t_all, nbr_all, eid_all, indptr_h, num_nodes
            kind = 'port'
        o.seed = 0
        nq = args.cpu_queries
        hn, ht = nodes[:nq].cpu().numpy(), times[:nq].cpu().numpy()
        cpu = {}
        for strat in ('recent', 'uniform', 'time_interval_aware'):
            o.sample_neighbor_strategy = strat
            o.random_state = np.random.RandomState(0)
            if strat == 'time_interval_aware':
                prob_h = samplers[strat].tia_prob[:base.num_half_edges].cpu().numpy()
                if kind == 'reference':
                    o.nodes_neighbor_sampled_probabilities = _NodeSlices(prob_h, indptr_h)
                else:
                    o.prob = prob_h
            n_ = nq if strat != 'time_interval_aware' else max(100, nq // 20)
            t1 = time.perf_counter()
            if strat == 'time_interval_aware':
                # RandomState.choice(p=float32 softmax) raises "probabilities do not sum to 1" on hubs with millions of
                # neighbours (utils/utils.py:183-187): such queries are counted, not timed
                done = raised = 0
                for q in range(n_):
                    try:
                        o.get_historical_neighbors(hn[q:q + 1], ht[q:q + 1], k)
                        done += 1
                    except ValueError:
                        raised += 1
                line['cpu_tia_reference_raises'] = raised
                n_ = max(done, 1)
            else:
                got = o.get_historical_neighbors(hn[:n_], ht[:n_], k)
            cpu[strat] = n_ / (time.perf_counter() - t1)
            if strat == 'recent':
                dv = samplers['recent'].get_historical_neighbors_device(nodes[:n_], times[:n_], k)
                line['parity_recent_bit_exact'] = bool(all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(got, dv)))
            if strat == 'uniform':
                # the replay mode consumes the same RandomState(0) stream from its start: same draws, same rows
                rr = replay_samplers['uniform']
                rr.reset_random_state()
                dv = rr.get_historical_neighbors_device(nodes[:n_], times[:n_], k)
                line['parity_uniform_replay_bit_exact'] = bool(all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(got, dv)))
        line['cpu_baseline'] = {'value': cpu['recent'], 'unit': 'queries/s', 'cores': 1, 'kind': kind,
                                'sample': f'{nq} queries (tia: {max(100, nq // 20)}), ' +
                                          ('unmodified reference query loop (utils/utils.py:130-214) over per-node slice views' if kind == 'reference'
                                           else 'oracle per-query python loop'), 'strategies': cpu}
    return line


class _NodeSlices:
    """``nodes_neighbor_ids``-style per-node arrays of the reference sampler as lazy slice views of one flat CSR array."""

    def __init__(self, flat, indptr):
        self.flat, self.indptr = flat, indptr

    def __getitem__(self, node):
        return self.flat[self.indptr[node]:self.indptr[node + 1]]

    def __len__(self):
        return len(self.indptr) - 1


class _LazyDeviceVector:
    """Host view of a device vector that copies only the slices asked for (the replay mode's per-query float32 softmax reads
    tia_prob[a:a+c] for the queried rows; the whole 1e8-event table would be a 1.6 GB copy)."""

    def __init__(self, vec):
        self.vec = vec

    def __getitem__(self, sl):
        return self.vec[sl].cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='dygformer_wiki',
                    choices=['dygformer_wiki', 'dygformer_lastfm', 'tgat_myket', 'tgn_reddit', 'sampler_sweep', 'tgat_train'])
    ap.add_argument('--batches-per-step', type=int, default=0)
    ap.add_argument('--cpu-batches', type=int, default=0)
    ap.add_argument('--events', type=int, default=100_000_000)
    ap.add_argument('--queries', type=int, default=1 << 24)
    ap.add_argument('--cpu-queries', type=int, default=20000)
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel directly instead of replaying the captured step')
    ap.add_argument('--no-eager', action='store_true', help='skip the eager-PyTorch-on-GPU run of the reference modules')
    ap.add_argument('--only-headline', action='store_true',
                    help='default workload only: do not append the `workloads` records of the other BASELINE configs')
    ap.add_argument('--save-dir', default='', help='also write the full JSON line of every extra workload into this directory')
    ap.add_argument('--ref-batch', type=int, default=0,
                    help='tgn_reddit only: events per sequential memory step instead of the reference default 200 (SURVEY 8d: B = 2,000 / '
                         '20,000 are different-semantics throughput points: a larger batch sees staler memories)')
    args = ap.parse_args()
    if args.ref_batch and args.ref_batch != REF_BATCH:
        if args.workload != 'tgn_reddit' or args.impl != 'ours':
            ap.error('--ref-batch applies to --workload tgn_reddit')
        globals()['REF_BATCH'] = args.ref_batch
    if args.impl == 'reference':
        if args.workload in ('sampler_sweep', 'tgat_train'):
            args.workload = 'dygformer_wiki'
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    ctx = Ctx()
    if args.workload == 'sampler_sweep':
        line = measure_sampler_sweep(ctx, args, args.steps, args.warmup)
    elif args.workload == 'tgat_train':
        line = measure_train(ctx, args, args.steps, args.warmup)
    else:
        line = run_ours(ctx, args)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.finish()


if __name__ == '__main__':
    main()
