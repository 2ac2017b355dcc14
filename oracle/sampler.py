"""Oracle restatement of ``NeighborSampler`` (reference ``utils/utils.py:71-302``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Storage differs from the reference (flat CSR arrays instead of python lists of
arrays) so 10^8-event graphs fit, but every query follows the reference's steps:
``np.searchsorted`` on the node's float64 times (``utils/utils.py:141``), then the
strategy-specific selection (``:176-209``).  Third-party arithmetic the reference
delegates to NumPy 2.3.5 / PyTorch 2.11.0 (``RandomState.choice``, ``argsort``,
``torch.softmax``) is delegated to the same libraries here.
"""
from __future__ import annotations

import numpy as np
import torch


class OracleSampler:
    def __init__(self, src, dst, eid, t, num_nodes=None, sample_neighbor_strategy='uniform',
                 time_scaling_factor=0.0, seed=None):
        """Follows ``get_neighbor_sampler`` (``utils/utils.py:283-302``): every event
        appends (dst, eid, t) to src's list and then (src, eid, t) to dst's list; each
        list is stable-sorted by time (``:100``)."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        eid = np.asarray(eid, dtype=np.int64)
        t = np.asarray(t, dtype=np.float64)
        if num_nodes is None:
            num_nodes = int(max(src.max(), dst.max())) + 1
        self.num_nodes = num_nodes
        self.sample_neighbor_strategy = sample_neighbor_strategy
        self.seed = seed
        E = len(src)
        owner = np.empty(2 * E, dtype=np.int64)
        owner[0::2] = src
        owner[1::2] = dst
        nbr = np.empty(2 * E, dtype=np.int64)
        nbr[0::2] = dst
        nbr[1::2] = src
        he_eid = np.repeat(eid, 2)
        he_t = np.repeat(t, 2)
        # stable by time, then stable by owner == per-owner stable sort by time
        o1 = np.argsort(he_t, kind='stable')
        o2 = np.argsort(owner[o1], kind='stable')
        order = o1[o2]
        self.nbr = nbr[order]
        self.eid = he_eid[order]
        self.t = he_t[order]
        counts = np.bincount(owner, minlength=num_nodes)
        self.indptr = np.zeros(num_nodes + 1, dtype=np.int64)
        np.cumsum(counts, out=self.indptr[1:])
        if sample_neighbor_strategy == 'time_interval_aware':
            self.time_scaling_factor = time_scaling_factor
            self.prob = self._sampled_probabilities()
        if seed is not None:
            self.random_state = np.random.RandomState(seed)

    def _sampled_probabilities(self):
        """``compute_sampled_probabilities`` (``utils/utils.py:112-128``) per node."""
        prob = np.empty(len(self.t), dtype=np.float64)
        nz = np.nonzero(np.diff(self.indptr))[0]
        with np.errstate(invalid='ignore', divide='ignore'):
            for v in nz:
                a, b = self.indptr[v], self.indptr[v + 1]
                tt = self.t[a:b] - np.max(self.t[a:b])
                e = np.exp(self.time_scaling_factor * tt)
                p = e / np.cumsum(e)
                p[np.isnan(p)] = -1e10
                prob[a:b] = p
        return prob

    def reset_random_state(self):
        self.random_state = np.random.RandomState(self.seed)

    def count_before(self, node_id, interact_time):
        """``find_neighbors_before`` (``utils/utils.py:141``): strict ``<`` via side='left'."""
        a, b = self.indptr[node_id], self.indptr[node_id + 1]
        return a, int(np.searchsorted(self.t[a:b], interact_time))

    def get_historical_neighbors(self, node_ids, node_interact_times, num_neighbors=20):
        """``get_historical_neighbors`` (``utils/utils.py:149-214``)."""
        assert num_neighbors > 0
        n, k = len(node_ids), num_neighbors
        out_n = np.zeros((n, k), dtype=np.int64)
        out_e = np.zeros((n, k), dtype=np.int64)
        out_t = np.zeros((n, k), dtype=np.float32)
        strat = self.sample_neighbor_strategy
        if strat not in ('uniform', 'recent', 'time_interval_aware'):
            raise ValueError(f'Not implemented error for sample_neighbor_strategy {strat}!')
        for idx, (v, tq) in enumerate(zip(node_ids, node_interact_times)):
            a, i = self.count_before(v, tq)
            if i == 0:
                continue
            if strat == 'recent':
                m = min(i, k)
                out_n[idx, k - m:] = self.nbr[a + i - m:a + i]
                out_e[idx, k - m:] = self.eid[a + i - m:a + i]
                out_t[idx, k - m:] = self.t[a + i - m:a + i]
            else:
                p = None
                if strat == 'time_interval_aware':
                    p = torch.softmax(torch.from_numpy(self.prob[a:a + i]).float(), dim=0).numpy()
                rs = self.random_state if self.seed is not None else np.random
                sel = rs.choice(a=i, size=k, p=p)
                out_n[idx] = self.nbr[a + sel]
                out_e[idx] = self.eid[a + sel]
                out_t[idx] = self.t[a + sel]
                pos = out_t[idx].argsort()
                out_n[idx] = out_n[idx][pos]
                out_e[idx] = out_e[idx][pos]
                out_t[idx] = out_t[idx][pos]
        return out_n, out_e, out_t

    def get_multi_hop_neighbors(self, num_hops, node_ids, node_interact_times, num_neighbors=20):
        """``get_multi_hop_neighbors`` (``utils/utils.py:216-252``)."""
        assert num_hops > 0
        n_, e_, t_ = self.get_historical_neighbors(node_ids, node_interact_times, num_neighbors)
        ln, le, lt = [n_], [e_], [t_]
        for _ in range(1, num_hops):
            n_, e_, t_ = self.get_historical_neighbors(ln[-1].flatten(), lt[-1].flatten(), num_neighbors)
            ln.append(n_.reshape(len(node_ids), -1))
            le.append(e_.reshape(len(node_ids), -1))
            lt.append(t_.reshape(len(node_ids), -1))
        return ln, le, lt

    def get_all_first_hop_neighbors(self, node_ids, node_interact_times):
        """``get_all_first_hop_neighbors`` (``utils/utils.py:254-273``)."""
        ln, le, lt = [], [], []
        for v, tq in zip(node_ids, node_interact_times):
            a, i = self.count_before(v, tq)
            ln.append(self.nbr[a:a + i])
            le.append(self.eid[a:a + i])
            lt.append(self.t[a:a + i])
        return ln, le, lt


def pad_sequences(node_ids, node_interact_times, ids_list, eids_list, times_list,
                  patch_size=1, max_input_sequence_length=256):
    """``DyGFormer.pad_sequences`` (``models/DyGFormer.py:196-245``): keep the most recent
    L-1 entries, prepend the node itself (edge 0, time = query time), right-pad with zeros
    to (batch max + 1) rounded up to a multiple of ``patch_size``."""
    assert max_input_sequence_length - 1 > 0
    L1 = max_input_sequence_length - 1
    ids_list = [x[-L1:] if len(x) > L1 else x for x in ids_list]
    eids_list = [x[-L1:] if len(x) > L1 else x for x in eids_list]
    times_list = [x[-L1:] if len(x) > L1 else x for x in times_list]
    max_len = max([len(x) for x in ids_list], default=0) + 1
    if max_len % patch_size != 0:
        max_len += patch_size - max_len % patch_size
    n = len(node_ids)
    pn = np.zeros((n, max_len), dtype=np.int64)
    pe = np.zeros((n, max_len), dtype=np.int64)
    pt = np.zeros((n, max_len), dtype=np.float32)
    for i in range(n):
        pn[i, 0] = node_ids[i]
        pt[i, 0] = node_interact_times[i]
        m = len(ids_list[i])
        if m > 0:
            pn[i, 1:m + 1] = ids_list[i]
            pe[i, 1:m + 1] = eids_list[i]
            pt[i, 1:m + 1] = times_list[i]
    return pn, pe, pt


def count_nodes_appearances(src_ids, dst_ids):
    """``NeighborCooccurrenceEncoder.count_nodes_appearances`` (``models/DyGFormer.py:337-393``).

    For every position of the src row: (occurrences of that id in the src row,
    occurrences in the dst row); likewise for the dst row; zero where id == 0.
    Returns float32 arrays (B, Ls, 2) and (B, Ld, 2)."""
    src_ids = np.asarray(src_ids)
    dst_ids = np.asarray(dst_ids)
    B = src_ids.shape[0]
    out_s = np.zeros(src_ids.shape + (2,), dtype=np.float32)
    out_d = np.zeros(dst_ids.shape + (2,), dtype=np.float32)
    for b in range(B):
        s, d = src_ids[b], dst_ids[b]
        su, sinv, sc = np.unique(s, return_inverse=True, return_counts=True)
        du, dinv, dc = np.unique(d, return_inverse=True, return_counts=True)
        smap = dict(zip(su.tolist(), sc.tolist()))
        dmap = dict(zip(du.tolist(), dc.tolist()))
        out_s[b, :, 0] = sc[sinv]
        out_s[b, :, 1] = [dmap.get(x, 0) for x in s.tolist()]
        out_d[b, :, 0] = [smap.get(x, 0) for x in d.tolist()]
        out_d[b, :, 1] = dc[dinv]
    out_s[src_ids == 0] = 0.0
    out_d[dst_ids == 0] = 0.0
    return out_s, out_d
