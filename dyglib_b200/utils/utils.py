"""Device-resident drop-in for the reference's ``NeighborSampler`` (``utils/utils.py:71-302``).

Same constructor, attributes and methods as the reference; the per-node python lists become one
CSR in HBM (16-byte half-edge records sorted per node by time) and every query method is a CUDA
kernel.  The numpy-in / numpy-out methods are the compatibility surface the reference's models
call; the ``*_device`` methods return CUDA tensors and are what the fused models use.

Random strategies (``uniform``, ``time_interval_aware``) consume the reference's RandomState stream
in call order (``rng='numpy_replay'``, bit-exact indices): prefix counts come back from the device,
the draws are replayed vectorised on the host with the same NumPy generator, and the gather / time
re-sort run on the device.  ``rng='philox'`` draws on the device (throughput mode, not bit-exact).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native, ops
from ..ops import _p, _stream


def set_random_seed(seed: int = 0):
    """``set_random_seed`` (``utils/utils.py:9-21``)."""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _as_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=_NP[dtype])).to(device, non_blocking=True)


_NP = {torch.int64: np.int64, torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32}


class NeighborSampler:

    def __init__(self, adj_list: list = None, sample_neighbor_strategy: str = 'uniform', time_scaling_factor: float = 0.0,
                 seed: int = None, device: str = 'cuda', rng: str = 'numpy_replay', tia_table: str = 'auto',
                 _edges=None):
        """
        :param adj_list: list of lists of (neighbor_id, edge_id, timestamp) tuples, as in the reference
        (``utils/utils.py:73-110``); position 0 is the empty list of the padding node.
        """
        ops.require_cuda()
        # an unknown strategy raises ValueError at query time, like the reference (utils/utils.py:211)
        self.sample_neighbor_strategy = sample_neighbor_strategy
        self.seed = seed
        self.device = torch.device(device)
        self.rng = rng
        self.use_fence = True
        if sample_neighbor_strategy == 'time_interval_aware':
            self.time_scaling_factor = time_scaling_factor
        if _edges is not None:
            src, dst, eid, t, num_nodes, presorted = _edges
            self._build(None, None, None, None, num_nodes, presorted, tia_table, events=(src, dst, eid, t))
        else:
            owner, nbr, eid, t, num_nodes, presorted = self._flatten_adj_list(adj_list)
            self._build(owner, nbr, eid, t, num_nodes, presorted, tia_table)
        if self.seed is not None:
            self.random_state = np.random.RandomState(self.seed)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _flatten_adj_list(adj_list):
        deg = np.fromiter((len(x) for x in adj_list), dtype=np.int64, count=len(adj_list))
        total = int(deg.sum())
        owner = np.repeat(np.arange(len(adj_list), dtype=np.int64), deg)
        nbr = np.empty(total, dtype=np.int64)
        eid = np.empty(total, dtype=np.int64)
        t = np.empty(total, dtype=np.float64)
        pos = 0
        for lst in adj_list:
            m = len(lst)
            if m:
                arr = np.asarray(lst, dtype=np.float64).reshape(m, 3)
                nbr[pos:pos + m] = arr[:, 0].astype(np.int64)
                eid[pos:pos + m] = arr[:, 1].astype(np.int64)
                t[pos:pos + m] = arr[:, 2]
                pos += m
        return owner, nbr, eid, t, len(adj_list), False

    def _build(self, owner, nbr, eid, t, num_nodes, presorted, tia_table, events=None):
        """Per-node time-sorted (stable) half-edge runs (``utils/utils.py:96-103``).

        ``events=(src, dst, eid, t)``: half-edge h = 2*e + side is owned by src[e] (side 0) or dst[e] (side 1),
        which is the reference's append order (``utils/utils.py:298-300``).  Otherwise ``owner/nbr/eid/t`` list
        the half-edges directly (adjacency-list constructor)."""
        dev = self.device
        lib = _native.load()
        self.num_nodes = int(num_nodes)
        if events is not None:
            src_d, dst_d, eid_d, t_d = (_as_dev(events[0], torch.int64, dev), _as_dev(events[1], torch.int64, dev),
                                        _as_dev(events[2], torch.int64, dev), _as_dev(events[3], torch.float64, dev))
            E = src_d.numel()
            n_half = 2 * E
            owner_d = torch.stack([src_d, dst_d], dim=1).reshape(-1)
            tkey = None if presorted else t_d.repeat_interleave(2)
            pack = (src_d, dst_d, eid_d, t_d)
            id_max = int(max(src_d.max().item(), dst_d.max().item(), eid_d.max().item())) if E else 0
        else:
            owner_d = _as_dev(owner, torch.int64, dev)
            nbr_d = _as_dev(nbr, torch.int64, dev)
            eid_d = _as_dev(eid, torch.int64, dev)
            t_d = _as_dev(t, torch.float64, dev)
            n_half = owner_d.numel()
            tkey = None if presorted else t_d
            pack = (nbr_d, nbr_d, eid_d, t_d)   # every half-edge is its own "event", always side 0
            id_max = int(max(nbr_d.max().item(), eid_d.max().item())) if n_half else 0
        if id_max >= 2 ** 31:
            raise ValueError('node / edge ids must fit in int32 for the device CSR')
        if n_half and (int(owner_d.max().item()) >= self.num_nodes or int(owner_d.min().item()) < 0):
            raise IndexError('node id out of range')
        # degrees -> indptr
        deg = torch.zeros(self.num_nodes, dtype=torch.int64, device=dev)
        if events is not None:
            _native.check(lib.dyg_csr_degrees(_p(pack[0]), _p(pack[1]), int(n_half // 2), self.num_nodes, _p(deg), _stream()))
            ops._count()
        elif n_half:
            deg = torch.bincount(owner_d, minlength=self.num_nodes)
        self.indptr = torch.zeros(self.num_nodes + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=self.indptr[1:])
        del deg
        # ordering: stable by time, then stable by owner — the package's own LSD radix sort (csrc/sort.cu, ops.stable_argsort)
        if n_half:
            owner32 = owner_d.to(torch.int32)
            if tkey is None:
                order = ops.stable_argsort(owner32)
            else:
                o1 = ops.stable_argsort(ops.float64_sort_key(tkey + 0.0))      # + 0.0: -0.0 -> +0.0, equal under the reference's comparison
                o2 = ops.stable_argsort(owner32[o1])
                order = o1[o2]
                del o1, o2
            del owner32
            if events is None:
                order = order * 2
        del owner_d, tkey
        self.halfedges = torch.empty((max(n_half, 1), 2), dtype=torch.float64, device=dev)  # 16 B per record
        if n_half:
            _native.check(lib.dyg_csr_pack(_p(order), _p(pack[0]), _p(pack[1]), _p(pack[2]), _p(pack[3]), int(n_half),
                                           _p(self.halfedges), _stream()))
            ops._count()
            del order
        self.num_half_edges = int(n_half)
        # fence index: last time of every complete 16^l-record block (dyg_csr_fence_build), 1/15 of a double per record
        n_f = int(lib.dyg_csr_fence_entries(int(n_half)))
        self.fence = torch.empty(n_f, dtype=torch.float64, device=dev) if n_f and self.use_fence else None
        if self.fence is not None:
            _native.check(lib.dyg_csr_fence_build(_p(self.halfedges), int(n_half), _p(self.fence), _stream()))
            ops._count()
        self._prob_host = None
        self.tia_cum = None
        self.tia_cum_fence = None
        self._philox_offset = 0
        if self.sample_neighbor_strategy == 'time_interval_aware':
            self._build_tia(tia_table)

    def _build_tia(self, tia_table):
        """``compute_sampled_probabilities`` (``utils/utils.py:112-128``)."""
        lib = _native.load()
        dev = self.device
        n_half = self.num_half_edges
        self.tia_cum = torch.empty(max(n_half, 1), dtype=torch.float64, device=dev)
        host = tia_table == 'host' or (tia_table == 'auto' and n_half <= 20_000_000)
        if host and n_half:
            # bit-exact table: same numpy float64 exp / cumsum per node as the reference
            indptr = self.indptr.cpu().numpy()
            t = self.halfedges[:n_half, 0].cpu().numpy()
            prob = np.empty(n_half, dtype=np.float64)
            with np.errstate(invalid='ignore', divide='ignore'):
                for v in np.nonzero(np.diff(indptr))[0]:
                    a, b = indptr[v], indptr[v + 1]
                    e = np.exp(self.time_scaling_factor * (t[a:b] - t[b - 1]))
                    p = e / np.cumsum(e)
                    p[np.isnan(p)] = -1e10
                    prob[a:b] = p
            self._prob_host = prob
            self.tia_prob = torch.from_numpy(prob).to(dev)
            _native.check(lib.dyg_csr_tia_cum(_p(self.tia_prob), _p(self.indptr), self.num_nodes, _p(self.tia_cum), _stream()))
        else:
            self.tia_prob = torch.empty(max(n_half, 1), dtype=torch.float64, device=dev)
            _native.check(lib.dyg_csr_tia_tables(_p(self.halfedges), _p(self.indptr), self.num_nodes,
                                                 float(self.time_scaling_factor), _p(self.tia_prob), _p(self.tia_cum), _stream()))
        ops._count()
        # fence index over the prefix table (16-ary levels, like the CSR fence) for the fused throughput kernel's CDF search
        n_f = int(lib.dyg_csr_fence_entries(int(n_half)))
        self.tia_cum_fence = torch.empty(n_f, dtype=torch.float64, device=dev) if n_f and self.use_fence else None
        if self.tia_cum_fence is not None:
            _native.check(lib.dyg_cum_fence_build(_p(self.tia_cum), int(n_half), _p(self.tia_cum_fence), _stream()))
            ops._count()

    # ------------------------------------------------------------------ reference-compatible attribute views
    def _node_slice(self, node_id):
        a = int(self.indptr[node_id].item())
        b = int(self.indptr[node_id + 1].item())
        return a, b

    def _host_records(self, a, b):
        rec = self.halfedges[a:b].cpu().numpy()
        t = rec[:, 0].copy()
        ints = rec[:, 1].copy().view(np.int32).reshape(-1, 2)
        return ints[:, 0].astype(np.int64), ints[:, 1].astype(np.int64), t

    def _host_lists(self):
        """Per-node host views of the device CSR (one D2H copy, cached): the reference's list attributes, built lazily."""
        if getattr(self, '_host_cache', None) is None:
            n = self.num_half_edges
            indptr = self.indptr.cpu().numpy()
            nbr, eid, t = self._host_records(0, n) if n else (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))
            cut = indptr[1:-1]
            cache = {'ids': np.split(nbr, cut), 'eids': np.split(eid, cut), 'times': np.split(t, cut)}
            if self.sample_neighbor_strategy == 'time_interval_aware':
                prob = self._prob_host if self._prob_host is not None else self.tia_prob[:n].cpu().numpy()
                cache['prob'] = np.split(prob, cut)
            self._host_cache = cache
        return self._host_cache

    @property
    def nodes_neighbor_ids(self):
        """``self.nodes_neighbor_ids`` of the reference (``utils/utils.py:85-103``): list over nodes of time-sorted neighbour ids."""
        return self._host_lists()['ids']

    @property
    def nodes_edge_ids(self):
        return self._host_lists()['eids']

    @property
    def nodes_neighbor_times(self):
        return self._host_lists()['times']

    @property
    def nodes_neighbor_sampled_probabilities(self):
        """``utils/utils.py:88-110``: only for ``time_interval_aware`` (AttributeError otherwise, as in the reference)."""
        if self.sample_neighbor_strategy != 'time_interval_aware':
            raise AttributeError("'NeighborSampler' object has no attribute 'nodes_neighbor_sampled_probabilities'")
        return self._host_lists()['prob']

    def find_neighbors_before(self, node_id: int, interact_time: float, return_sampled_probabilities: bool = False):
        """``find_neighbors_before`` (``utils/utils.py:130-147``); host views of one node's prefix."""
        a, b = self._node_slice(int(node_id))
        nbr, eid, t = self._host_records(a, b)
        i = int(np.searchsorted(t, interact_time))
        prob = None
        if return_sampled_probabilities:
            prob = self.tia_prob[a:a + i].cpu().numpy()
        return nbr[:i], eid[:i], t[:i], prob

    def reset_random_state(self):
        """``reset_random_state`` (``utils/utils.py:275-280``)."""
        self.random_state = np.random.RandomState(self.seed)
        self._philox_offset = 0

    # ------------------------------------------------------------------ device queries
    def _queries(self, node_ids, node_interact_times):
        ids = _as_dev(node_ids, torch.int64, self.device)
        # float32 query times (hop >= 2, models/TGAT.py:108) are promoted exactly, as np.searchsorted does
        tq = _as_dev(node_interact_times, torch.float64, self.device) if not (
            isinstance(node_interact_times, torch.Tensor) and node_interact_times.dtype == torch.float64
            and node_interact_times.device == self.device) else node_interact_times.contiguous()
        return ids, tq

    def count_before_device(self, ids, tq):
        cnt = torch.empty(ids.numel(), dtype=torch.int32, device=self.device)
        _native.check(_native.load().dyg_count_before(_p(self.halfedges), _p(self.indptr), self.num_nodes, _p(self.fence), self.num_half_edges, _p(ids), _p(tq),
                                                      ids.numel(), _p(cnt), _stream()))
        ops._count()
        return cnt

    def get_historical_neighbors_device(self, node_ids, node_interact_times, num_neighbors: int = 20):
        """Device version of ``get_historical_neighbors`` (``utils/utils.py:149-214``):
        returns CUDA tensors (int64 (n,k), int64 (n,k), float32 (n,k))."""
        assert num_neighbors > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        strat = self.sample_neighbor_strategy
        if strat not in ('uniform', 'recent', 'time_interval_aware'):
            raise ValueError(f'Not implemented error for sample_neighbor_strategy {strat}!')
        ids, tq = self._queries(node_ids, node_interact_times)
        n, k = ids.numel(), int(num_neighbors)
        dev = self.device
        out_n = torch.empty((n, k), dtype=torch.int64, device=dev)
        out_e = torch.empty((n, k), dtype=torch.int64, device=dev)
        out_t = torch.empty((n, k), dtype=torch.float32, device=dev)
        lib = _native.load()
        if strat == 'recent':
            _native.check(lib.dyg_sample_recent(_p(self.halfedges), _p(self.indptr), self.num_nodes, _p(self.fence), self.num_half_edges, _p(ids), _p(tq), n, k,
                                                _p(out_n), _p(out_e), _p(out_t), None, _stream()))
            ops._count()
            return out_n, out_e, out_t
        if self.rng == 'philox':
            # throughput mode: one fused kernel, counter-based draws (not the reference's RandomState stream)
            off = self._philox_offset
            _native.check(lib.dyg_sample_random(_p(self.halfedges), _p(self.indptr), self.num_nodes, _p(self.fence), self.num_half_edges,
                                                _p(self.tia_cum) if strat == 'time_interval_aware' else None,
                                                _p(self.tia_cum_fence) if strat == 'time_interval_aware' else None,
                                                _p(ids), _p(tq), n, k, int(self.seed or 0), int(off),
                                                _p(out_n), _p(out_e), _p(out_t), _stream()))
            self._philox_offset = off + n * k
            ops._count()
            return out_n, out_e, out_t
        cnt = self.count_before_device(ids, tq)
        sel = self._draw(ids, cnt, n, k)
        _native.check(lib.dyg_sample_indexed(_p(self.halfedges), _p(self.indptr), _p(ids), _p(cnt), _p(sel), n, k,
                                             _p(out_n), _p(out_e), _p(out_t), _stream()))
        ops._count()
        return out_n, out_e, out_t

    def _draw(self, ids, cnt, n, k):
        """Positions in [0, cnt) for every (query, draw), consuming the RNG stream like the reference."""
        lib = _native.load()
        dev = self.device
        strat = self.sample_neighbor_strategy
        sel = torch.empty((n, k), dtype=torch.int64, device=dev)
        if self.rng == 'philox':
            u = torch.empty(n * k, dtype=torch.float64, device=dev)
            off = getattr(self, '_philox_offset', 0)
            _native.check(lib.dyg_philox_uniform(int(self.seed or 0), int(off), n * k, _p(u), _stream()))
            self._philox_offset = off + ((n * k + 1) // 2) * 2
            if strat == 'uniform':
                _native.check(lib.dyg_draw_uniform(_p(cnt), _p(u), n, k, _p(sel), _stream()))
            else:
                _native.check(lib.dyg_draw_tia(_p(self.tia_cum), _p(self.indptr), _p(ids), _p(cnt), _p(u), n, k, _p(sel), _stream()))
            ops._count(2)
            return sel
        # numpy_replay: the reference draws per query, in order, only for queries with a non-empty prefix
        cnt_h = cnt.cpu().numpy().astype(np.int64)
        rs = self.random_state if self.seed is not None else np.random
        nz = np.nonzero(cnt_h)[0]
        sel_h = np.zeros((n, k), dtype=np.int64)
        if strat == 'uniform':
            if len(nz):
                # RandomState.choice(a=c, size=k) == randint(0, c, k): one vectorised call replays the same
                # masked-rejection word stream as the per-query calls
                sel_h[nz] = rs.randint(0, np.repeat(cnt_h[nz], k)).reshape(len(nz), k)
        else:
            if self._prob_host is None:
                self._prob_host = self.tia_prob[:self.num_half_edges].cpu().numpy()
            indptr_h = self.indptr[ids].cpu().numpy()
            for q in nz:  # O(prefix) float32 softmax per query, as the reference does (utils/utils.py:183)
                a, c = indptr_h[q], cnt_h[q]
                p = torch.softmax(torch.from_numpy(self._prob_host[a:a + c]).float(), dim=0).numpy()
                sel_h[q] = rs.choice(a=c, size=k, p=p)
        sel.copy_(torch.from_numpy(sel_h), non_blocking=False)
        # The reference re-sorts a query's draws with ndarray.argsort() on their float32 times (utils/utils.py:196), an UNSTABLE sort:
        # distinct records whose times round to the same float32 come out in numpy's order, not in draw order.  Parity mode replays
        # that call on the host (times of the drawn records come back from the device) and hands the kernel positions that are
        # already in the reference's order; its own stable rank sort is then the identity.
        if len(nz):
            a0 = self.indptr[ids].reshape(n, 1)
            t32 = self.halfedges[:, 0][(a0 + sel).clamp_(max=max(self.num_half_edges - 1, 0))].float().cpu().numpy()
            order = np.argsort(t32[nz], axis=1)
            sel_h[nz] = np.take_along_axis(sel_h[nz], order, axis=1)
            sel.copy_(torch.from_numpy(sel_h), non_blocking=False)
        return sel

    def get_all_first_hop_neighbors_device(self, node_ids, node_interact_times, max_input_sequence_length: int,
                                           patch_size: int = 1, group_size: int = 0):
        """Fused ``get_all_first_hop_neighbors`` + ``DyGFormer.pad_sequences`` (``utils/utils.py:254-273``,
        ``models/DyGFormer.py:196-245``) truncated to the most recent L-1 entries.
        Returns padded (n, W) int64 ids, int64 edge ids, float32 times with W = L rounded up to patch_size,
        per-row lengths (int32) and, if group_size > 0, per-group max lengths (int32)."""
        L = int(max_input_sequence_length)
        assert L - 1 > 0, 'Maximal number of neighbors for each node should be greater than 1!'
        ids, tq = self._queries(node_ids, node_interact_times)
        n = ids.numel()
        W = ((L + patch_size - 1) // patch_size) * patch_size
        dev = self.device
        pn = torch.empty((n, W), dtype=torch.int64, device=dev)
        pe = torch.empty((n, W), dtype=torch.int64, device=dev)
        pt = torch.empty((n, W), dtype=torch.float32, device=dev)
        ln = torch.empty(n, dtype=torch.int32, device=dev)
        gmax = None
        if group_size > 0:
            gmax = torch.zeros((n + group_size - 1) // group_size, dtype=torch.int32, device=dev)
        _native.check(_native.load().dyg_first_hop_pad(_p(self.halfedges), _p(self.indptr), self.num_nodes, _p(self.fence), self.num_half_edges, _p(ids), _p(tq), n,
                                                       L, W, _p(pn), _p(pe), _p(pt), _p(ln), _p(gmax), int(group_size), _stream()))
        ops._count()
        return pn, pe, pt, ln, gmax

    # ------------------------------------------------------------------ reference numpy API
    def get_historical_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20):
        """``get_historical_neighbors`` (``utils/utils.py:149-214``): numpy in, numpy out."""
        a, b, c = self.get_historical_neighbors_device(node_ids, node_interact_times, num_neighbors)
        return a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()

    def get_multi_hop_neighbors(self, num_hops: int, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20):
        """``get_multi_hop_neighbors`` (``utils/utils.py:216-252``)."""
        assert num_hops > 0, 'Number of sampled hops should be greater than 0!'
        n = len(node_ids)
        a, b, c = self.get_historical_neighbors_device(node_ids, node_interact_times, num_neighbors)
        ln, le, lt = [a], [b], [c]
        for _ in range(1, num_hops):
            # hop h queries the previous hop's neighbours at their float32 interaction times
            a, b, c = self.get_historical_neighbors_device(ln[-1].reshape(-1), lt[-1].reshape(-1).double(), num_neighbors)
            ln.append(a.reshape(n, -1))
            le.append(b.reshape(n, -1))
            lt.append(c.reshape(n, -1))
        return ([x.cpu().numpy() for x in ln], [x.cpu().numpy() for x in le], [x.cpu().numpy() for x in lt])

    def get_all_first_hop_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray):
        """``get_all_first_hop_neighbors`` (``utils/utils.py:254-273``): three lists of ragged arrays
        (int64 ids, int64 edge ids, float64 times), one entry per query, full history."""
        ids, tq = self._queries(node_ids, node_interact_times)
        cnt = self.count_before_device(ids, tq).cpu().numpy()
        start = self.indptr[ids].cpu().numpy()
        ln, le, lt = [], [], []
        for a, c in zip(start, cnt):
            nb, ei, t = self._host_records(int(a), int(a + c))
            ln.append(nb)
            le.append(ei)
            lt.append(t)
        return ln, le, lt


def get_neighbor_sampler(data, sample_neighbor_strategy: str = 'uniform', time_scaling_factor: float = 0.0, seed: int = None,
                         device: str = 'cuda', rng: str = 'numpy_replay', tia_table: str = 'auto'):
    """``get_neighbor_sampler`` (``utils/utils.py:283-302``): ``data`` has ``src_node_ids``, ``dst_node_ids``,
    ``edge_ids``, ``node_interact_times``.  Every event contributes (dst, eid, t) to src's list and then
    (src, eid, t) to dst's list; the interleaved half-edge order reproduces the reference's append order."""
    src = np.asarray(data.src_node_ids, dtype=np.int64)
    dst = np.asarray(data.dst_node_ids, dtype=np.int64)
    eid = np.asarray(data.edge_ids, dtype=np.int64)
    t = np.asarray(data.node_interact_times, dtype=np.float64)
    E = len(src)
    num_nodes = int(max(src.max(), dst.max())) + 1 if E else 1
    presorted = bool(E == 0 or np.all(t[1:] >= t[:-1]))
    return NeighborSampler(None, sample_neighbor_strategy, time_scaling_factor, seed, device, rng, tia_table,
                           _edges=(src, dst, eid, t, num_nodes, presorted))


class NegativeEdgeSampler(object):
    """``NegativeEdgeSampler`` (``utils/utils.py:303-390``), ``random`` strategy: the negative sources / destinations of a batch are
    drawn uniformly from the unique sources / destinations with the reference's ``RandomState(seed).randint`` stream (global
    ``np.random`` without a seed), so a seeded sampler returns the reference's ids.  ``sample_device`` returns them as CUDA
    tensors for the fused models.  The ``historical`` / ``inductive`` strategies enumerate python sets of edge tuples whose
    iteration order is an implementation detail of the reference; they are outside this path and raise."""

    def __init__(self, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, interact_times: np.ndarray = None,
                 last_observed_time: float = None, negative_sample_strategy: str = 'random', seed: int = None):
        self.seed = seed
        self.negative_sample_strategy = negative_sample_strategy
        self.src_node_ids = src_node_ids
        self.dst_node_ids = dst_node_ids
        self.interact_times = interact_times
        self.unique_src_node_ids = np.unique(src_node_ids)
        self.unique_dst_node_ids = np.unique(dst_node_ids)
        self.last_observed_time = last_observed_time
        if self.seed is not None:
            self.random_state = np.random.RandomState(self.seed)

    def sample(self, size: int, batch_src_node_ids: np.ndarray = None, batch_dst_node_ids: np.ndarray = None,
               current_batch_start_time: float = 0.0, current_batch_end_time: float = 0.0):
        """``sample`` (``utils/utils.py:349-376``)."""
        if self.negative_sample_strategy == 'random':
            return self.random_sample(size=size)
        if self.negative_sample_strategy in ('historical', 'inductive'):
            raise NotImplementedError(f'negative_sample_strategy {self.negative_sample_strategy}: use the reference sampler (host-side set logic)')
        raise ValueError(f'Not implemented error for negative_sample_strategy {self.negative_sample_strategy}!')

    def random_sample(self, size: int):
        """``random_sample`` (``utils/utils.py:378-390``): source draw first, then destination draw."""
        rs = np.random if self.seed is None else self.random_state
        si = rs.randint(0, len(self.unique_src_node_ids), size)
        di = rs.randint(0, len(self.unique_dst_node_ids), size)
        return self.unique_src_node_ids[si], self.unique_dst_node_ids[di]

    def sample_device(self, size: int, device='cuda'):
        """``random_sample`` as int64 CUDA tensors (same stream, same ids)."""
        s, d = self.random_sample(size)
        return torch.from_numpy(np.ascontiguousarray(s)).to(device), torch.from_numpy(np.ascontiguousarray(d)).to(device)

    def reset_random_state(self):
        """``reset_random_state`` (``utils/utils.py:472-477``)."""
        self.random_state = np.random.RandomState(self.seed)
