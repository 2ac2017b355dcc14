"""Generate tests/golden/tgat_train.npz, memory_train.npz and dygformer_train.npz: loss and every parameter gradient of one training step
of the UNMODIFIED reference TGAT / MemoryModel (TGN, DyRep, JODIE) / DyGFormer + MergeLayer link predictor (train mode, dropout 0 so that
the step is deterministic; the memory models first advance 8 batches without gradients), /root/reference, build container only:
    python scripts/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')

from helpers import (small_graph, deterministic_state_dict, tgat_train_step, memory_train_step, dygformer_train_step,  # noqa: E402
                     DYG_TRAIN_CASES, compact_grads)
from utils.utils import get_neighbor_sampler  # noqa: E402  (reference)
from utils.DataLoader import Data  # noqa: E402
from models.TGAT import TGAT  # noqa: E402
from models.MemoryModel import MemoryModel  # noqa: E402
from models.DyGFormer import DyGFormer  # noqa: E402
from models.modules import MergeLayer  # noqa: E402


def main():
    g = small_graph(seed=11)
    data = Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    m = TGAT(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(data, 'recent'), 100, 2, 2, 0.0).train()
    m.load_state_dict(deterministic_state_dict(m.state_dict(), 1))
    pred = MergeLayer(172, 172, 172, 1).train()
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    params = {'model.' + k: v for k, v in m.named_parameters()}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    out = tgat_train_step(lambda s, d, t, k: m.compute_src_dst_node_temporal_embeddings(s, d, t, k), lambda a, b: pred(a, b), params)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'tgat_train.npz'), **compact_grads(out))
    print('loss', out['loss'], {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith('grad.')})
    g = small_graph(seed=13)
    data = Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    allout = {}
    for name in ('TGN', 'DyRep', 'JODIE'):
        m = MemoryModel(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(data, 'recent'), 100, name, num_layers=1,
                        num_heads=2, dropout=0.0, src_node_mean_time_shift=3.0, src_node_std_time_shift=50.0,
                        dst_node_mean_time_shift_dst=5.0, dst_node_std_time_shift=70.0).train()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), 3))
        m.memory_bank.__init_memory_bank__()
        pred = MergeLayer(172, 172, 172, 1).train()
        pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
        params = {'model.' + k: v for k, v in m.named_parameters() if v.requires_grad}
        params.update({'pred.' + k: v for k, v in pred.named_parameters()})
        out = memory_train_step(m, lambda a, b: pred(a, b), params)
        print(name, 'loss', out['loss'], {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith('grad.')})
        allout.update({name + '.' + k: v for k, v in out.items()})
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'memory_train.npz'), **compact_grads(allout))
    g = small_graph(seed=12)
    data = Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    allout = {}
    for P, L in DYG_TRAIN_CASES:
        m = DyGFormer(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(data, 'recent'), 100, 50, patch_size=P, num_layers=2,
                      num_heads=2, dropout=0.0, max_input_sequence_length=L).train()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), 2))
        pred = MergeLayer(172, 172, 172, 1).train()
        pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
        params = {'model.' + k: v for k, v in m.named_parameters()}
        params.update({'pred.' + k: v for k, v in pred.named_parameters()})
        out = dygformer_train_step(m, lambda a, b: pred(a, b), params)
        print('DyGFormer', P, L, 'loss', out['loss'], {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith('grad.')})
        allout.update({f'P{P}_L{L}.' + k: v for k, v in out.items()})
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'dygformer_train.npz'), **compact_grads(allout))


if __name__ == '__main__':
    main()
