"""Generate tests/golden/tcl.npz from the UNMODIFIED reference TCL (/root/reference, build container only): eval-mode
embeddings of two batches and the loss / parameter gradients of one training step (dropout 0):
    python scripts/make_golden_tcl.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, '/root/reference')

from helpers import small_graph, deterministic_state_dict, run_tcl_cases, tcl_train_step, compact_grads  # noqa: E402
from utils.utils import get_neighbor_sampler  # noqa: E402  (reference)
from utils.DataLoader import Data  # noqa: E402
from models.TCL import TCL  # noqa: E402
from models.modules import MergeLayer  # noqa: E402


def main():
    g = small_graph(seed=15)
    data = Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)

    def make(dropout, train):
        m = TCL(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(data, 'recent'), 100, num_layers=2, num_heads=2,
                num_depths=21, dropout=dropout)
        m = m.train() if train else m.eval()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), 7))
        return m
    out = run_tcl_cases(make(0.1, False))
    m = make(0.0, True)
    pred = MergeLayer(172, 172, 172, 1).train()
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    params = {'model.' + k: v for k, v in m.named_parameters()}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    tr = tcl_train_step(m, lambda a, b: pred(a, b), params)
    out.update({'train.' + k: v for k, v in tr.items()})
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'tcl.npz'), **compact_grads(out))
    print('loss', tr['loss'], {k: float(np.abs(v).max()) for k, v in out.items()})


if __name__ == '__main__':
    main()
