"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the reference tree does not travel to the GPU box):
    python scripts/make_golden.py
Inputs are regenerated from seeds by tests/helpers.py, so the fixtures hold outputs only.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')

from helpers import small_graph, deterministic_state_dict, batches, make_queries  # noqa: E402
from utils.utils import get_neighbor_sampler  # noqa: E402  (reference)
from utils.DataLoader import Data  # noqa: E402
from models.TGAT import TGAT  # noqa: E402
from models.DyGFormer import DyGFormer  # noqa: E402
from models.MemoryModel import MemoryModel  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def ref_sampler(g, strategy, seed=None, tsf=0.0):
    data = Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    return get_neighbor_sampler(data, strategy, time_scaling_factor=tsf, seed=seed)


def golden_sampler():
    g = small_graph(seed=7)
    out = {}
    for strategy, seed in (('recent', None), ('uniform', 3), ('time_interval_aware', 3)):
        s = ref_sampler(g, strategy, seed, 1e-5)
        rng = np.random.default_rng(0)
        for k, f32 in ((20, False), (3, True), (1, False)):
            nodes, times = make_queries(g, 400, rng, f32)
            a, b, c = s.get_historical_neighbors(nodes, times, k)
            out[f'{strategy}_k{k}_nbr'], out[f'{strategy}_k{k}_eid'], out[f'{strategy}_k{k}_t'] = a, b, c
    s = ref_sampler(g, 'recent')
    rng = np.random.default_rng(5)
    nodes, times = make_queries(g, 100, rng)
    ln, le, lt = s.get_multi_hop_neighbors(2, nodes, times, 3)
    for h in range(2):
        out[f'multihop_{h}_nbr'], out[f'multihop_{h}_eid'], out[f'multihop_{h}_t'] = ln[h], le[h], lt[h]
    # first hop + pad + co-occurrence
    rng = np.random.default_rng(1)
    n1, t1 = make_queries(g, 200, rng)
    n2, _ = make_queries(g, 200, rng)
    dyg = DyGFormer(g.node_raw_features, g.edge_raw_features, s, 100, 50, patch_size=4, max_input_sequence_length=32)
    pads = []
    for nodes in (n1, n2):
        a = s.get_all_first_hop_neighbors(nodes, t1)
        out.setdefault('firsthop_len', np.array([len(x) for x in a[0]]))
        pads.append(dyg.pad_sequences(nodes, t1, list(a[0]), list(a[1]), list(a[2]), 4, 32))
    for i, p in enumerate(pads):
        out[f'pad{i}_nbr'], out[f'pad{i}_eid'], out[f'pad{i}_t'] = p
    cs, cd = dyg.neighbor_co_occurrence_encoder.count_nodes_appearances(pads[0][0], pads[1][0])
    out['cooc_src'], out['cooc_dst'] = cs.numpy(), cd.numpy()
    np.savez_compressed(os.path.join(OUT, 'sampler.npz'), **out)


def golden_models():
    out = {}
    with torch.no_grad():
        g = small_graph(seed=11)
        m = TGAT(g.node_raw_features, g.edge_raw_features, ref_sampler(g, 'recent'), 100, 2, 2, 0.1).eval()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), 1))
        for bi, (src, dst, t, _, neg) in enumerate(batches(g, 2000, 2, 40)):
            for tag, d in (('pos', dst), ('neg', neg)):
                a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                out[f'tgat_{bi}_{tag}_src'], out[f'tgat_{bi}_{tag}_dst'] = a.numpy(), b.numpy()
        g = small_graph(seed=12)
        for P, L in ((2, 16), (1, 8), (4, 32)):
            m = DyGFormer(g.node_raw_features, g.edge_raw_features, ref_sampler(g, 'recent'), 100, 50, patch_size=P, num_layers=2,
                          num_heads=2, dropout=0.1, max_input_sequence_length=L).eval()
            m.load_state_dict(deterministic_state_dict(m.state_dict(), 2))
            for bi, (src, dst, t, _, neg) in enumerate(batches(g, 1000, 2, 50)):
                for tag, d in (('pos', dst), ('neg', neg)):
                    a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t)
                    out[f'dygformer_P{P}_L{L}_{bi}_{tag}_src'], out[f'dygformer_P{P}_L{L}_{bi}_{tag}_dst'] = a.numpy(), b.numpy()
        g = small_graph(seed=13)
        for name in ('TGN', 'DyRep', 'JODIE'):
            m = MemoryModel(g.node_raw_features, g.edge_raw_features, ref_sampler(g, 'recent'), 100, name, num_layers=1, num_heads=2,
                            dropout=0.1, src_node_mean_time_shift=3.0, src_node_std_time_shift=50.0,
                            dst_node_mean_time_shift_dst=5.0, dst_node_std_time_shift=70.0).eval()
            m.load_state_dict(deterministic_state_dict(m.state_dict(), 3))
            for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 12, 30)):
                ra = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
                rb = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
                if bi >= 9:
                    out[f'{name}_{bi}_neg_src'], out[f'{name}_{bi}_neg_dst'] = ra[0].numpy(), ra[1].numpy()
                    out[f'{name}_{bi}_pos_src'], out[f'{name}_{bi}_pos_dst'] = rb[0].numpy(), rb[1].numpy()
            out[f'{name}_memory'] = m.memory_bank.node_memories.data.numpy()
            out[f'{name}_last_update'] = m.memory_bank.node_last_updated_times.data.numpy()
    np.savez_compressed(os.path.join(OUT, 'models.npz'), **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    golden_sampler()
    golden_models()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
