"""Seeded synthetic temporal graphs shaped like the datasets BASELINE.json names.

The reference ships no data (``/root/reference/.MISSING_LARGE_BLOBS``), so every
config runs on graphs made here.  Conventions follow the reference's
preprocessing (``preprocess_data/preprocess_data.py:76-79,101-108``): node 0 and
edge 0 are padding, feature row 0 is all zeros, users are 1..Nu, items are
Nu+1..Nu+Ni, edge ids are 1..E in chronological order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class TemporalGraph:
    """Same attribute names as the reference's ``Data`` (``utils/DataLoader.py:46-64``)."""
    src_node_ids: np.ndarray          # int64 (E,)
    dst_node_ids: np.ndarray          # int64 (E,)
    node_interact_times: np.ndarray   # float64 (E,)
    edge_ids: np.ndarray              # int64 (E,)
    labels: np.ndarray                # float64 (E,) (unused by link prediction)
    num_nodes: int                    # includes padding node 0
    node_raw_features: np.ndarray | None = None   # float32 (num_nodes, F)
    edge_raw_features: np.ndarray | None = None   # float32 (E + 1, F)

    @property
    def num_interactions(self) -> int:
        return len(self.src_node_ids)


def _zipf_ranks(rng: np.random.Generator, n_items: int, alpha: float, size: int) -> np.ndarray:
    """Draw ``size`` ranks in [0, n_items) with p(rank) proportional to (rank+1)^-alpha."""
    w = np.arange(1, n_items + 1, dtype=np.float64) ** (-alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    out = np.empty(size, dtype=np.int64)
    chunk = 1 << 24
    for s in range(0, size, chunk):
        u = rng.random(min(chunk, size - s))
        out[s:s + len(u)] = np.searchsorted(cdf, u, side='right')
    np.minimum(out, n_items - 1, out=out)
    return out


def make_graph(num_events: int, num_users: int, num_items: int, t_max: float, seed: int,
               feat_dim: int = 172, with_features: bool = True,
               user_alpha: float = 0.8, item_alpha: float = 1.0) -> TemporalGraph:
    rng = np.random.default_rng(seed)
    user_perm = rng.permutation(num_users)
    item_perm = rng.permutation(num_items)
    src = 1 + user_perm[_zipf_ranks(rng, num_users, user_alpha, num_events)]
    dst = 1 + num_users + item_perm[_zipf_ranks(rng, num_items, item_alpha, num_events)]
    t_max_i = int(t_max)
    if t_max_i + 1 >= 4 * num_events and num_events <= (1 << 25):
        t = np.sort(rng.choice(t_max_i + 1, size=num_events, replace=False)).astype(np.float64)
    else:
        # strictly increasing integers: random gaps >= 1 scaled to end near t_max
        mean_gap = max(1.0, t_max_i / max(1, num_events))
        gaps = 1 + rng.integers(0, max(1, int(2 * mean_gap) - 1), size=num_events)
        t = np.cumsum(gaps).astype(np.float64)
    eid = np.arange(1, num_events + 1, dtype=np.int64)
    num_nodes = num_users + num_items + 1
    node_feat = edge_feat = None
    if with_features:
        node_feat = (rng.standard_normal((num_nodes, feat_dim), dtype=np.float32) * np.float32(0.1))
        edge_feat = (rng.standard_normal((num_events + 1, feat_dim), dtype=np.float32) * np.float32(0.1))
        node_feat[0] = 0.0
        edge_feat[0] = 0.0
    return TemporalGraph(src.astype(np.int64), dst.astype(np.int64), t, eid,
                         np.zeros(num_events), num_nodes, node_feat, edge_feat)


# the five BASELINE.json configs.
CONFIGS = {
    'tgat_myket':     dict(num_events=694_121,     num_users=10_000,    num_items=7_988,     t_max=1.7e7,     seed=1),
    'dygformer_wiki': dict(num_events=157_474,     num_users=8_227,     num_items=1_000,     t_max=2_678_373, seed=2),
    'tgn_reddit':     dict(num_events=672_447,     num_users=10_000,    num_items=984,       t_max=2_678_390, seed=3),
    'dygformer_lastfm': dict(num_events=1_293_103, num_users=980,       num_items=1_000,     t_max=1.37e8,    seed=4),
    'sampler_100m':   dict(num_events=100_000_000, num_users=8_000_000, num_items=2_000_000, t_max=3.0e8,     seed=5,
                           with_features=False),
}


def make_config_graph(name: str, scale: float = 1.0, **overrides) -> TemporalGraph:
    """Build one of the named graphs; ``scale`` < 1 shrinks events and nodes for tests."""
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    if scale != 1.0:
        cfg['num_events'] = max(16, int(cfg['num_events'] * scale))
        cfg['num_users'] = max(4, int(cfg['num_users'] * scale))
        cfg['num_items'] = max(4, int(cfg['num_items'] * scale))
    return make_graph(**cfg)


def random_negative_dst(graph: TemporalGraph, size: int, rng: np.random.RandomState) -> np.ndarray:
    """Random negative destinations, the ``random`` mode of the reference's
    NegativeEdgeSampler (``utils/utils.py:378-390``): uniform index into the unique dst ids."""
    uniq = np.unique(graph.dst_node_ids)
    return uniq[rng.randint(0, len(uniq), size)]
