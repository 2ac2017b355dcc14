"""GPU parity: device sampler / pad / co-occurrence kernels vs the oracle and the reference's golden vectors.
Integer, index and time outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from helpers import (run_sampler_cases, oracle_factories, cuda_factories, load_golden, small_graph, make_queries)
from dyglib_b200.synthetic import make_graph, make_config_graph
from oracle.sampler import OracleSampler, pad_sequences, count_nodes_appearances

pytestmark = pytest.mark.gpu


def cuda_pad(s, g, nodes, times, lists, P, L):
    pn, pe, pt, ln, _ = s.get_all_first_hop_neighbors_device(nodes, times, L, P)
    Lp = (int(ln.max().item()) + P - 1) // P * P
    return pn[:, :Lp].cpu().numpy(), pe[:, :Lp].cpu().numpy(), pt[:, :Lp].cpu().numpy()


def cuda_cooc(a, b):
    from dyglib_b200.models.DyGFormer import NeighborCooccurrenceEncoder
    enc = NeighborCooccurrenceEncoder(50, 'cuda').to('cuda')
    x, y = enc.count_nodes_appearances(a, b)
    return x.cpu().numpy(), y.cpu().numpy()


def assert_same(got, want):
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype, (k, got[k].dtype, want[k].dtype)
        assert got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), k


def test_sampler_cases_match_golden_and_oracle():
    csampler = cuda_factories()[0]
    got = run_sampler_cases(csampler, cuda_pad, cuda_cooc)
    assert_same(got, load_golden('sampler.npz'))
    osampler = oracle_factories()[0]
    want = run_sampler_cases(osampler, lambda s, g, n, t, lists, P, L: pad_sequences(n, t, lists[0], lists[1], lists[2], P, L),
                             count_nodes_appearances)
    assert_same(got, want)


def _pair(g, strategy='recent', seed=None, tsf=0.0, **kw):
    from dyglib_b200.utils.utils import get_neighbor_sampler
    return (get_neighbor_sampler(g, strategy, tsf, seed, **kw),
            OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, strategy, tsf, seed))


def _check(c, o, nodes, times, k):
    a = c.get_historical_neighbors(nodes, times, k)
    b = o.get_historical_neighbors(nodes, times, k)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x, y)


@pytest.mark.parametrize('k', [1, 2, 5, 10, 20, 33, 64, 130])
def test_recent_all_lane_widths(k):
    g = small_graph(seed=21, E=6000, nu=30, ni=10)
    c, o = _pair(g)
    rng = np.random.default_rng(k)
    nodes, times = make_queries(g, 700, rng)
    _check(c, o, nodes, times, k)


def test_recent_edge_cases():
    """zero history, deg < k, query time equal to an event time (strict <), float32-rounded times > 2^24,
    node 0, duplicate timestamps, self loops, unsorted input."""
    rng = np.random.default_rng(3)
    E = 4000
    src = rng.integers(1, 40, E)
    dst = rng.integers(1, 40, E)                       # non-bipartite, self loops occur
    t = np.sort(rng.integers(1, 500, E)).astype(np.float64) + 1.0e8   # many duplicates, > 2^24
    g = make_graph(16, 4, 4, 100, 0, with_features=False)
    g.src_node_ids, g.dst_node_ids, g.node_interact_times = src, dst, t
    g.edge_ids = np.arange(1, E + 1)
    g.num_nodes = 40
    c, o = _pair(g)
    nodes = rng.integers(0, 40, 3000)
    times = t[rng.integers(0, E, 3000)]
    times[:500] = times[:500].astype(np.float32)        # hop-2 style float32-rounded query times
    times[500:600] = 0.0                                 # before everything
    times[600:700] = 1e12                                # after everything
    for k in (20, 4):
        _check(c, o, nodes, times, k)
        _check(c, o, nodes, times.astype(np.float32), k)
    # unsorted input order: both must stable-sort by time per node
    perm = rng.permutation(E)
    g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids = src[perm], dst[perm], t[perm], g.edge_ids[perm]
    c, o = _pair(g)
    _check(c, o, nodes, times, 7)


def test_adj_list_constructor_matches():
    from dyglib_b200.utils.utils import NeighborSampler
    g = small_graph(seed=5, E=1500)
    adj = [[] for _ in range(g.num_nodes)]
    for s, d, e, t in zip(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times):
        adj[s].append((d, e, t))
        adj[d].append((s, e, t))
    c = NeighborSampler(adj, 'recent')
    o = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
    nodes, times = make_queries(g, 500, np.random.default_rng(0))
    _check(c, o, nodes, times, 20)
    nb, ei, tt, _ = c.find_neighbors_before(int(nodes[100]), float(times[100]))
    a, i = o.count_before(nodes[100], times[100])
    assert np.array_equal(nb, o.nbr[a:a + i]) and np.array_equal(ei, o.eid[a:a + i]) and np.array_equal(tt, o.t[a:a + i])
    # the reference's list attributes (utils/utils.py:85-110) as lazy host views of the device CSR
    assert len(c.nodes_neighbor_ids) == len(adj) and len(c.nodes_neighbor_ids[0]) == 0
    vq = int(g.src_node_ids[7])
    for v in (1, vq, g.num_nodes - 1):
        ref = sorted(adj[v], key=lambda x: x[2])
        assert np.array_equal(c.nodes_neighbor_ids[v], np.array([x[0] for x in ref], dtype=np.int64))
        assert np.array_equal(c.nodes_edge_ids[v], np.array([x[1] for x in ref], dtype=np.int64))
        assert np.array_equal(c.nodes_neighbor_times[v], np.array([x[2] for x in ref], dtype=np.float64))
    with pytest.raises(AttributeError):
        c.nodes_neighbor_sampled_probabilities
    ct = NeighborSampler(adj, 'time_interval_aware', time_scaling_factor=1e-5, seed=0)
    tv = np.array([x[2] for x in sorted(adj[vq], key=lambda x: x[2])])
    e = np.exp(1e-5 * (tv - tv.max()))
    assert np.allclose(ct.nodes_neighbor_sampled_probabilities[vq], e / np.cumsum(e), rtol=1e-12)


@pytest.mark.parametrize('strategy', ['uniform', 'time_interval_aware'])
def test_random_strategies_replay_stream_across_calls(strategy):
    """The RNG stream is consumed in call order, also across recursion-style repeated calls."""
    g = small_graph(seed=31, E=5000, nu=50, ni=20)
    c, o = _pair(g, strategy, seed=11, tsf=2e-5)
    rng = np.random.default_rng(9)
    for k in (20, 5):
        nodes, times = make_queries(g, 300, rng)
        _check(c, o, nodes, times, k)
    c.reset_random_state()
    o.reset_random_state()
    nodes, times = make_queries(g, 200, rng)
    a = c.get_multi_hop_neighbors(2, nodes, times, 4)
    b = o.get_multi_hop_neighbors(2, nodes, times, 4)
    for la, lb in zip(a, b):
        for x, y in zip(la, lb):
            assert np.array_equal(x, y)


@pytest.mark.parametrize('strategy', ['uniform', 'time_interval_aware'])
def test_replay_mode_reproduces_numpy_tie_order(strategy):
    """Timestamps above 2^24: distinct records round to the same float32 time, and the reference's unstable argsort
    (utils/utils.py:196) decides their order.  Replay mode must come out in that order, not in draw order."""
    g = small_graph(seed=77, E=6000, nu=30, ni=12)
    g.node_interact_times = g.node_interact_times + 3.0e8        # float32 spacing 32: neighbouring events collide
    c, o = _pair(g, strategy, seed=3, tsf=1e-6)
    rng = np.random.default_rng(4)
    nodes, times = make_queries(g, 400, rng)
    a = c.get_historical_neighbors(nodes, times, 20)
    b = o.get_historical_neighbors(nodes, times, 20)
    ties = sum(len(np.unique(r)) < len(r) for r in b[2])
    assert ties > 100                                             # the case is exercised
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_tia_device_table_mismatch_rate():
    """Device prefix-CDF draw fed with the replayed random_sample stream: indices equal the reference's unless a
    uniform lands within float32-softmax rounding of a CDF boundary (expected rate <~ 1e-6 per draw)."""
    g = small_graph(seed=41, E=8000, nu=40, ni=15)
    c, o = _pair(g, 'time_interval_aware', seed=5, tsf=1e-5)
    from dyglib_b200 import _native
    from dyglib_b200.ops import _p, _stream
    rng = np.random.default_rng(2)
    nodes, times = make_queries(g, 3000, rng)
    k = 20
    want = o.get_historical_neighbors(nodes, times, k)
    ids, tq = c._queries(nodes, times)
    cnt = c.count_before_device(ids, tq)
    cnt_h = cnt.cpu().numpy()
    rs = np.random.RandomState(5)
    u = np.zeros((len(nodes), k))
    for q in np.nonzero(cnt_h)[0]:
        u[q] = rs.random_sample(k)
    u_d = torch.from_numpy(u).cuda()
    sel = torch.empty((len(nodes), k), dtype=torch.int64, device='cuda')
    _native.check(_native.load().dyg_draw_tia(_p(c.tia_cum), _p(c.indptr), _p(ids), _p(cnt), _p(u_d), len(nodes), k, _p(sel), _stream()))
    out = [torch.empty((len(nodes), k), dtype=d, device='cuda') for d in (torch.int64, torch.int64, torch.float32)]
    _native.check(_native.load().dyg_sample_indexed(_p(c.halfedges), _p(c.indptr), _p(ids), _p(cnt), _p(sel), len(nodes), k,
                                                    _p(out[0]), _p(out[1]), _p(out[2]), _stream()))
    mism = float((out[1].cpu().numpy() != want[1]).mean())
    assert mism <= 1e-4, mism


def test_philox_mode_draws_valid_history():
    g = small_graph(seed=51, E=5000)
    from dyglib_b200.utils.utils import get_neighbor_sampler
    for strategy in ('uniform', 'time_interval_aware'):
        c = get_neighbor_sampler(g, strategy, 1e-5, seed=7, rng='philox')
        nodes, times = make_queries(g, 2000, np.random.default_rng(4))
        nb, ei, tt = c.get_historical_neighbors(nodes, times, 20)
        has = nb[:, 0] != 0
        assert has.any()
        assert (np.diff(tt[has], axis=1) >= 0).all()                      # rows re-sorted by time
        assert (tt[has] < times[has, None].astype(np.float32) + 1).all()  # only history
        # every sampled edge id is a real interaction of that node
        e = ei[has].reshape(-1)
        owner = np.repeat(nodes[has], 20)
        assert ((g.src_node_ids[e - 1] == owner) | (g.dst_node_ids[e - 1] == owner)).all()


def test_large_random_recent_and_firsthop():
    """>= 1e5 random queries on a wikipedia-sized graph against the oracle."""
    g = make_config_graph('dygformer_wiki', with_features=False)
    c, o = _pair(g)
    rng = np.random.default_rng(0)
    nodes, times = make_queries(g, 100_000, rng)
    _check(c, o, nodes, times, 20)
    nodes, times = nodes[:3000], times[:3000]
    for P, L in ((2, 64), (16, 512)):
        lists = o.get_all_first_hop_neighbors(nodes, times)
        want = pad_sequences(nodes, times, lists[0], lists[1], lists[2], P, L)
        got = cuda_pad(c, g, nodes, times, None, P, L)
        for x, y in zip(got, want):
            assert x.dtype == y.dtype and np.array_equal(x, y)
    ln, le, lt = c.get_all_first_hop_neighbors(nodes[:50], times[:50])
    ol = o.get_all_first_hop_neighbors(nodes[:50], times[:50])
    for i in range(50):
        assert np.array_equal(ln[i], ol[0][i]) and np.array_equal(le[i], ol[1][i]) and np.array_equal(lt[i], ol[2][i])


@pytest.mark.parametrize('Ls,Ld', [(64, 64), (512, 512), (8, 40), (1, 1)])
def test_cooc_counts_exact(Ls, Ld):
    rng = np.random.default_rng(Ls + Ld)
    B = 37
    s = rng.integers(0, 12, (B, Ls))      # few distinct ids -> many collisions, zeros = padding
    d = rng.integers(0, 12, (B, Ld))
    s[0] = 0
    d[1] = 0
    want = count_nodes_appearances(s, d)
    got = cuda_cooc(s, d)
    for x, y in zip(got, want):
        assert np.array_equal(x, y)


def test_full_size_properties():
    """BASELINE-size (1.29 M events, LastFM-shaped) properties that need no oracle: sortedness, strictness,
    left padding, idempotence, agreement between count_before and the sampled tail."""
    g = make_config_graph('dygformer_lastfm', with_features=False)
    from dyglib_b200.utils.utils import get_neighbor_sampler
    c = get_neighbor_sampler(g, 'recent')
    rng = np.random.default_rng(1)
    e = rng.integers(0, g.num_interactions, 1 << 18)
    nodes = np.where(rng.integers(0, 2, len(e)) == 1, g.src_node_ids[e], g.dst_node_ids[e])
    times = g.node_interact_times[e]
    nb, ei, tt = c.get_historical_neighbors(nodes, times, 20)
    nb2, ei2, tt2 = c.get_historical_neighbors(nodes, times, 20)
    assert np.array_equal(nb, nb2) and np.array_equal(ei, ei2) and np.array_equal(tt, tt2)
    valid = nb != 0
    assert (np.diff(valid.astype(np.int8), axis=1) >= 0).all()           # zeros only on the left
    assert (np.diff(np.where(valid, tt, np.float32(-1.0)), axis=1) >= 0).all()   # ascending times
    assert (g.node_interact_times[ei[valid] - 1] < np.repeat(times[:, None], 20, 1)[valid]).all()   # strict history
    ids, tq = c._queries(nodes, times)
    cnt = c.count_before_device(ids, tq).cpu().numpy()
    assert np.array_equal(np.minimum(cnt, 20), valid.sum(1))
    deg = np.bincount(np.concatenate([g.src_node_ids, g.dst_node_ids]), minlength=g.num_nodes)
    assert (cnt <= deg[nodes]).all()


@pytest.mark.parametrize('dups,index64', [(False, False), (True, False), (True, True)])
def test_fence_index_matches_searchsorted(dups, index64, monkeypatch):
    """The fenced lower bound (dyg_csr_fence_build + level descent) equals np.searchsorted(side='left') on every node's
    run, for hubs spanning 1-5 fence levels, runs that start / end inside a 16-record block, duplicate timestamps, and
    query times equal to / between / outside the stored times; the un-indexed search gives the same counts."""
    from dyglib_b200.utils.utils import NeighborSampler
    if index64:
        monkeypatch.setenv('DYG_FENCE_INDEX64', '1')      # int64 index arithmetic (CSR of 2^31 half-edges and more)
    rng = np.random.default_rng(11 + dups)
    degs = [0, 1, 15, 16, 17, 31, 47, 48, 49, 63, 255, 256, 257, 700, 4095, 4097, 70001, 3, 1048590 // 4, 5]
    owner = np.repeat(np.arange(1, len(degs) + 1), degs)
    n_half = len(owner)
    t = (rng.integers(0, 3000, n_half) if dups else rng.permutation(n_half * 2)[:n_half]).astype(np.float64)
    dev = torch.device('cuda')
    # half-edge h owned by owner[h] at time t[h] (adjacency-list constructor path: unsorted input, stable sort)
    s = object.__new__(NeighborSampler)
    s.device, s.use_fence, s.sample_neighbor_strategy, s.seed, s.rng = dev, True, 'recent', None, 'numpy_replay'
    s._build(owner, rng.integers(1, 100, n_half), np.arange(1, n_half + 1), t, len(degs) + 1, False, 'auto')
    assert s.fence is not None
    indptr = s.indptr.cpu().numpy()
    rec_t = s.halfedges[:n_half, 0].cpu().numpy()
    nq = 20000
    nodes = rng.integers(0, len(degs) + 1, nq)
    times = np.empty(nq)
    for i, v in enumerate(nodes):
        a, b = indptr[v], indptr[v + 1]
        mode = i % 4
        if b == a or mode == 0:
            times[i] = rng.integers(-5, 2 * n_half + 5)
        elif mode == 1:
            times[i] = rec_t[rng.integers(a, b)]            # equal to a stored time: strict <
        elif mode == 2:
            times[i] = rec_t[rng.integers(a, b)] + 0.5
        else:
            times[i] = rec_t[b - 1] + rng.integers(0, 2)     # at / past the end
    want = np.array([np.searchsorted(rec_t[indptr[v]:indptr[v + 1]], tq, side='left') for v, tq in zip(nodes, times)], dtype=np.int32)
    ids, tq = s._queries(nodes, times)
    got = s.count_before_device(ids, tq).cpu().numpy()
    assert np.array_equal(got, want)
    fence, s.fence = s.fence, None
    assert np.array_equal(s.count_before_device(ids, tq).cpu().numpy(), want)
    s.fence = fence
    for k in (3, 10, 20, 50, 100):                          # every lane width of sample_recent + the first-hop kernel
        a = s.get_historical_neighbors_device(ids, tq, k)
        s.fence = None
        b = s.get_historical_neighbors_device(ids, tq, k)
        s.fence = fence
        assert all(torch.equal(x, y) for x, y in zip(a, b))
        last = a[2][:, -1].cpu().numpy()
        idx = indptr[nodes] + want - 1
        assert np.array_equal(last[want > 0], rec_t[idx[want > 0]].astype(np.float32))
    a = s.get_all_first_hop_neighbors_device(ids, tq, 64, 2)
    s.fence = None
    b = s.get_all_first_hop_neighbors_device(ids, tq, 64, 2)
    assert all(torch.equal(x, y) for x, y in zip(a[:4], b[:4]))


@pytest.mark.parametrize('index64,tsf', [(False, 3e-5), (True, 3e-5), (False, 0.5)])
def test_tia_cum_fence_matches_binary_search_and_numpy(index64, tsf, monkeypatch):
    """Fused throughput kernel, time_interval_aware: the CDF search through the fence index over the prefix table returns the
    same draws as the plain binary search (fence disabled), and both equal np.searchsorted(cum[:cnt], u * cum[cnt-1], 'right')
    with the kernel's own Philox uniforms.  Degrees cross every block / level boundary of the 16-ary index; the large time
    scaling factor produces runs of equal table entries (the secant step must not divide by a zero increment)."""
    if index64:
        monkeypatch.setenv('DYG_FENCE_INDEX64', '1')
    from dyglib_b200 import _native
    from dyglib_b200.ops import _p, _stream
    from dyglib_b200.utils.utils import NeighborSampler
    rng = np.random.default_rng(23)
    degs = [0, 1, 15, 16, 17, 31, 33, 47, 48, 49, 63, 255, 256, 257, 511, 513, 700, 4095, 4097, 70001, 3, 300000, 5]
    owner = np.repeat(np.arange(1, len(degs) + 1), degs)
    n_half = len(owner)
    t = rng.permutation(n_half * 2)[:n_half].astype(np.float64)
    dev = torch.device('cuda')
    s = object.__new__(NeighborSampler)
    s.device, s.use_fence, s.sample_neighbor_strategy, s.seed, s.rng = dev, True, 'time_interval_aware', 9, 'philox'
    s.time_scaling_factor = tsf     # 0.5: exp underflows for most of a run -> -1e10 rows, a flat (tied) start of the prefix table
    s._build(owner, rng.integers(1, 100, n_half), np.arange(1, n_half + 1), t, len(degs) + 1, False, 'device')
    assert s.tia_cum_fence is not None
    indptr = s.indptr.cpu().numpy()
    rec_t = s.halfedges[:n_half, 0].cpu().numpy()
    cum = s.tia_cum[:n_half].cpu().numpy()
    nq, k = 6000, 20
    nodes = rng.integers(0, len(degs) + 1, nq)
    times = np.empty(nq)
    for i, v in enumerate(nodes):
        a, b = indptr[v], indptr[v + 1]
        times[i] = rng.integers(-5, 2 * n_half + 5) if (b == a or i % 3 == 0) else rec_t[rng.integers(a, b)] + (i % 2)
    ids, tq = s._queries(nodes, times)
    s._philox_offset = 0
    with_fence = s.get_historical_neighbors_device(ids, tq, k)
    fence, s.tia_cum_fence = s.tia_cum_fence, None
    s._philox_offset = 0
    without = s.get_historical_neighbors_device(ids, tq, k)
    s.tia_cum_fence = fence
    assert all(torch.equal(x, y) for x, y in zip(with_fence, without))
    # numpy restatement from the same uniforms
    # draw (q, j) of the fused kernel = first double of Philox counter q * k + j = element 2 (q k + j) of dyg_philox_uniform
    u = torch.empty(2 * nq * k, dtype=torch.float64, device=dev)
    _native.check(_native.load().dyg_philox_uniform(9, 0, 2 * nq * k, _p(u), _stream()))
    u = u.cpu().numpy()[::2].reshape(nq, k)
    cnt = s.count_before_device(ids, tq).cpu().numpy()
    rec = s.halfedges[:n_half].cpu().numpy()
    eids = rec[:, 1].copy().view(np.int32).reshape(-1, 2)[:, 1]
    got_e = with_fence[1].cpu().numpy()
    for q in range(nq):
        a, c = indptr[nodes[q]], cnt[q]
        if c == 0:
            assert (got_e[q] == 0).all()
            continue
        tot = cum[a + c - 1]
        sel = np.searchsorted(cum[a:a + c], u[q] * tot, side='right') if tot > 0 else np.floor(u[q] * c).astype(np.int64)
        sel = np.minimum(sel, c - 1)
        want_t = rec_t[a + sel].astype(np.float32)
        order = np.lexsort((np.arange(k), want_t))
        assert np.array_equal(got_e[q], eids[a + sel][order].astype(np.int64)), q


@pytest.mark.parametrize('strategy', ['uniform', 'time_interval_aware'])
def test_philox_mode_follows_the_reference_sampling_law(strategy):
    """The Philox throughput modes do not replay the reference's MT19937 stream, but they must draw from the same distribution
    (utils/utils.py:176-199): index i of the prefix with probability 1 / cnt ('uniform') or softmax_f32(prob[:cnt])_i
    ('time_interval_aware', prob from compute_sampled_probabilities, :112-128).  Chi-square test of the index histogram of 80,000
    draws per hub query against those probabilities."""
    from scipy import stats
    from dyglib_b200.utils.utils import get_neighbor_sampler
    tsf = 2e-5
    g = small_graph(seed=53, E=6000, nu=40, ni=15)
    c = get_neighbor_sampler(g, strategy, tsf, seed=11, rng='philox')
    o = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, strategy, tsf, 11)
    deg = np.bincount(np.concatenate([g.src_node_ids, g.dst_node_ids]), minlength=g.num_nodes)
    hubs = np.argsort(-deg)[[0, 3, 12]]
    reps, k = 4000, 20
    for v in hubs:
        tq = float(np.quantile(g.node_interact_times, 0.8)) + 0.5
        a, cnt = o.count_before(int(v), tq)
        eid = o.eid[a:a + cnt]
        assert cnt >= 30
        if strategy == 'uniform':
            p = np.full(cnt, 1.0 / cnt)
        else:
            p = torch.softmax(torch.from_numpy(o.prob[a:a + cnt]).float(), dim=0).double().numpy()
            p /= p.sum()
        _, ei, _ = c.get_historical_neighbors(np.full(reps, v), np.full(reps, tq), k)
        pos = {int(e): i for i, e in enumerate(eid)}            # edge ids are unique per interaction
        idx = np.array([pos[int(e)] for e in ei.reshape(-1)])
        obs = np.bincount(idx, minlength=cnt).astype(np.float64)
        exp = p * obs.sum()
        # merge the bins whose expectation is too small for the chi-square approximation
        small = exp < 5
        if small.any():
            obs = np.concatenate([obs[~small], [obs[small].sum()]])
            exp = np.concatenate([exp[~small], [exp[small].sum()]])
        stat, pval = stats.chisquare(obs, exp)
        assert pval > 1e-4, (strategy, int(v), cnt, stat, pval)
        # and the test has power: a law shifted towards recent entries is rejected
        wrong = p * np.linspace(0.7, 1.3, cnt)
        wrong /= wrong.sum()
        we = wrong * obs.sum() if not small.any() else None
        if we is not None:
            assert stats.chisquare(obs, we)[1] < 1e-6


@pytest.mark.parametrize('n,bits', [(1, 3), (31, 2), (4096, 8), (4097, 9), (100_003, 5), (1_000_003, 20), (300_000, 31)])
def test_own_stable_radix_sort_matches_torch_stable_sort(n, bits):
    """csrc/sort.cu (the CSR build's ordering): positions equal torch.sort(stable=True) on heavily tied keys, 4- and 8-byte keys,
    and float64 times through the order-preserving bit pattern."""
    import torch
    from dyglib_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(n + bits)
    k32 = torch.randint(0, 2 ** bits, (n,), device='cuda', generator=g, dtype=torch.int64).to(torch.int32)
    assert torch.equal(ops.stable_argsort(k32), torch.sort(k32.to(torch.int64), stable=True).indices)
    k64 = torch.randint(0, 2 ** 62, (n,), device='cuda', generator=g, dtype=torch.int64) >> (62 - min(bits + 30, 62))
    assert torch.equal(ops.stable_argsort(k64), torch.sort(k64, stable=True).indices)
    t = (torch.randint(0, 2 ** bits, (n,), device='cuda', generator=g).double() - 2 ** (bits - 1)) * 0.37
    assert torch.equal(ops.stable_argsort(ops.float64_sort_key(t + 0.0)), torch.sort(t, stable=True).indices)
