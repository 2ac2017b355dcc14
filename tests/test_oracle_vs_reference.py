"""Pins the oracle against the unmodified reference, run side by side (build container only)."""
import numpy as np
import pytest
import torch

from dyglib_b200.synthetic import make_graph
from oracle.sampler import OracleSampler, pad_sequences, count_nodes_appearances
from oracle.models import OracleTGAT, OracleDyGFormer, OracleMemoryModel, OracleGraphMixer, OracleTCL

pytestmark = pytest.mark.reference


def small_graph(seed=7, E=3000, nu=60, ni=25, tmax=200000.0, F=172):
    return make_graph(E, nu, ni, tmax, seed, feat_dim=F)


def ref_sampler(ref, g, strategy, seed=None, tsf=0.0):
    data = ref['DataLoader'].Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    return ref['utils'].get_neighbor_sampler(data, strategy, time_scaling_factor=tsf, seed=seed)


def oracle_sampler(g, strategy, seed=None, tsf=0.0):
    return OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes,
                         strategy, tsf, seed)


def queries(g, n, rng, with_f32=False):
    e = rng.integers(0, g.num_interactions, n)
    side = rng.integers(0, 2, n).astype(bool)
    nodes = np.where(side, g.src_node_ids[e], g.dst_node_ids[e])
    times = g.node_interact_times[e].copy()
    nodes[: n // 20] = 0                       # padding-node queries
    times[n // 20: n // 10] += 0.5             # strictly between events
    if with_f32:
        times = times.astype(np.float32)
    return nodes, times


@pytest.mark.parametrize('strategy', ['recent', 'uniform', 'time_interval_aware'])
def test_sampler_matches_reference(ref, strategy):
    g = small_graph()
    seed = None if strategy == 'recent' else 3
    rs, os_ = ref_sampler(ref, g, strategy, seed, 1e-5), oracle_sampler(g, strategy, seed, 1e-5)
    rng = np.random.default_rng(0)
    for k, f32 in ((20, False), (3, True), (1, False)):
        nodes, times = queries(g, 500, rng, f32)
        a = rs.get_historical_neighbors(nodes, times, k)
        b = os_.get_historical_neighbors(nodes, times, k)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)


def test_first_hop_pad_cooc_match_reference(ref):
    g = small_graph(seed=9)
    rs, os_ = ref_sampler(ref, g, 'recent'), oracle_sampler(g, 'recent')
    rng = np.random.default_rng(1)
    nodes, times = queries(g, 200, rng)
    a = rs.get_all_first_hop_neighbors(nodes, times)
    b = os_.get_all_first_hop_neighbors(nodes, times)
    for la, lb in zip(a, b):
        assert all(np.array_equal(x, y) for x, y in zip(la, lb))
    dyg = ref['DyGFormer'].DyGFormer(g.node_raw_features, g.edge_raw_features, rs, 100, 50, patch_size=4,
                                     max_input_sequence_length=32)
    pa = dyg.pad_sequences(nodes, times, list(a[0]), list(a[1]), list(a[2]), 4, 32)
    pb = pad_sequences(nodes, times, b[0], b[1], b[2], 4, 32)
    for x, y in zip(pa, pb):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    nodes2, _ = queries(g, 200, rng)
    a2 = rs.get_all_first_hop_neighbors(nodes2, times)
    pa2 = dyg.pad_sequences(nodes2, times, list(a2[0]), list(a2[1]), list(a2[2]), 4, 32)
    ca = dyg.neighbor_co_occurrence_encoder.count_nodes_appearances(pa[0], pa2[0])
    cb = count_nodes_appearances(pa[0], pa2[0])
    for x, y in zip(ca, cb):
        assert np.array_equal(x.numpy(), y)


def _batches(g, start, nb, B, seed=0):
    rng = np.random.RandomState(seed)
    uniq = np.unique(g.dst_node_ids)
    for b in range(nb):
        sl = slice(start + b * B, start + (b + 1) * B)
        neg = uniq[rng.randint(0, len(uniq), B)]
        yield g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], g.edge_ids[sl], neg


def test_tgat_matches_reference(ref):
    g = small_graph(seed=11)
    torch.manual_seed(0)
    m = ref['TGAT'].TGAT(g.node_raw_features, g.edge_raw_features, ref_sampler(ref, g, 'recent'), 100, 2, 2, 0.1).eval()
    o = OracleTGAT(m.state_dict(), g.node_raw_features, g.edge_raw_features, oracle_sampler(g, 'recent'), 2, 2)
    with torch.no_grad():
        for src, dst, t, _, neg in _batches(g, 2000, 2, 40):
            for d in (dst, neg):
                ra = m.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                oa = o.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                for x, y in zip(ra, oa):
                    torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('P,L', [(2, 16), (1, 8)])
def test_dygformer_matches_reference(ref, P, L):
    g = small_graph(seed=12)
    torch.manual_seed(0)
    m = ref['DyGFormer'].DyGFormer(g.node_raw_features, g.edge_raw_features, ref_sampler(ref, g, 'recent'), 100, 50,
                                   patch_size=P, num_layers=2, num_heads=2, dropout=0.1,
                                   max_input_sequence_length=L).eval()
    o = OracleDyGFormer(m.state_dict(), g.node_raw_features, g.edge_raw_features, oracle_sampler(g, 'recent'), 50, P, 2, 2, L)
    with torch.no_grad():
        for src, dst, t, _, neg in _batches(g, 1000, 2, 50):
            for d in (dst, neg):
                ra = m.compute_src_dst_node_temporal_embeddings(src, d, t)
                oa = o.compute_src_dst_node_temporal_embeddings(src, d, t)
                for x, y in zip(ra, oa):
                    torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('name', ['TGN', 'DyRep', 'JODIE'])
def test_memory_model_matches_reference(ref, name):
    g = small_graph(seed=13)
    torch.manual_seed(0)
    m = ref['MemoryModel'].MemoryModel(g.node_raw_features, g.edge_raw_features, ref_sampler(ref, g, 'recent'), 100, name,
                                       num_layers=1, num_heads=2, dropout=0.1, src_node_mean_time_shift=3.0,
                                       src_node_std_time_shift=50.0, dst_node_mean_time_shift_dst=5.0,
                                       dst_node_std_time_shift=70.0).eval()
    o = OracleMemoryModel(m.state_dict(), g.node_raw_features, g.edge_raw_features, oracle_sampler(g, 'recent'), name, 1, 2,
                          3.0, 50.0, 5.0, 70.0)
    with torch.no_grad():
        for src, dst, t, eid, neg in _batches(g, 0, 12, 30):
            ra = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            oa = o.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            rb = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            ob = o.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            for x, y in zip(ra + rb, oa + ob):
                torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.memory_bank.node_memories.data, o.memory, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.memory_bank.node_last_updated_times.data, o.last_update)


def test_graphmixer_matches_reference(ref):
    """The reference's own default-initialised weights (not the fixture's), uniform sampling for the link encoder's neighbours:
    the oracle must follow the reference's RandomState stream through both sampler calls of every root set."""
    g = small_graph(seed=21)
    torch.manual_seed(1)
    m = ref['GraphMixer'].GraphMixer(g.node_raw_features, g.edge_raw_features, ref_sampler(ref, g, 'uniform', seed=4), 100, num_tokens=10,
                                     num_layers=2, dropout=0.1).eval()
    o = OracleGraphMixer(m.state_dict(), g.node_raw_features, g.edge_raw_features, oracle_sampler(g, 'uniform', seed=4), 2)
    with torch.no_grad():
        for src, dst, t, _, neg in _batches(g, 1500, 2, 30):
            for d in (dst, neg):
                ra = m.compute_src_dst_node_temporal_embeddings(src, d, t, 10, 50)
                oa = o.compute_src_dst_node_temporal_embeddings(src, d, t, 10, 50)
                for x, y in zip(ra, oa):
                    torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)


def test_tcl_matches_reference(ref):
    g = small_graph(seed=22)
    torch.manual_seed(2)
    m = ref['TCL'].TCL(g.node_raw_features, g.edge_raw_features, ref_sampler(ref, g, 'recent'), 100, num_layers=2, num_heads=2,
                       num_depths=11, dropout=0.1).eval()
    o = OracleTCL(m.state_dict(), g.node_raw_features, g.edge_raw_features, oracle_sampler(g, 'recent'), 2, 2)
    with torch.no_grad():
        for src, dst, t, _, neg in _batches(g, 1500, 2, 30):
            for d in (dst, neg):
                ra = m.compute_src_dst_node_temporal_embeddings(src, d, t, 10)
                oa = o.compute_src_dst_node_temporal_embeddings(src, d, t, 10)
                for x, y in zip(ra, oa):
                    torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6)
