"""GPU parity: fused CUDA modules / models vs the oracle and the reference's golden vectors.
Tolerance: north-star 1e-3 relative in fp32 (asserted as rtol 1e-3 + atol 2e-4 on O(0.1..1) values); the
tighter observed error is asserted where the path is pure fp32 FFMA."""
import numpy as np
import pytest
import torch

from helpers import (run_model_cases, oracle_factories, cuda_factories, load_golden, small_graph, batches,
                     deterministic_state_dict)
from dyglib_b200.synthetic import make_config_graph
from oracle import models as om

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 2e-4


def close(a, b, msg='', rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=msg)


def test_time_encoder_matches_torch_cpu():
    from dyglib_b200.models.modules import TimeEncoder
    enc = TimeEncoder(100).cuda()
    with torch.no_grad():
        enc.w.bias.copy_(0.3 * torch.randn(100, generator=torch.Generator().manual_seed(0)))
    g = torch.Generator().manual_seed(1)
    dt = torch.cat([torch.rand(64, 40, generator=g) * 3e6, torch.rand(8, 40, generator=g) * 1.4e8,
                    torch.zeros(1, 40), -torch.rand(2, 40, generator=g) * 1e5])
    sd = {'w.weight': enc.w.weight.detach().cpu(), 'w.bias': enc.w.bias.detach().cpu()}
    want = om.time_encode(sd, '', dt)
    got = enc(dt.cuda())
    assert got.shape == want.shape
    # the fp32 argument fma(dt, w, b) is bit-identical; only the cos evaluation differs (< 3e-7 absolute)
    close(got, want, rtol=0, atol=5e-7)


@pytest.mark.parametrize('k', [20, 10, 3])
def test_multi_head_attention_module(k):
    from dyglib_b200.models.modules import MultiHeadAttention
    F_, E_, T_ = 172, 172, 100
    m = MultiHeadAttention(F_, E_, T_, 2, 0.1).cuda().eval()
    sd = deterministic_state_dict(m.state_dict(), 5)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(k)
    n = 97
    x = torch.randn(n, F_, generator=g)
    tq = torch.randn(n, 1, T_, generator=g)
    nf = torch.randn(n, k, F_, generator=g)
    tf = torch.randn(n, k, T_, generator=g)
    ef = torch.randn(n, k, E_, generator=g)
    ids = torch.randint(0, 5, (n, k), generator=g).numpy()
    ids[0] = 0                                              # fully masked row -> uniform attention (-1e10 fill)
    ids[1, : k - 1] = 0
    want_o, want_s = om.temporal_attention(sd, '', x, tq, nf, tf, ef, ids, 2)
    got_o, got_s = m(x.cuda(), tq.cuda(), nf.cuda(), tf.cuda(), ef.cuda(), ids)
    close(got_s, want_s, 'scores')
    close(got_o, want_o, 'output')
    assert abs(float(got_s[0].sum()) - 2.0) < 1e-4


def test_merge_layer_and_transformer_encoder():
    from dyglib_b200.models.modules import MergeLayer
    from dyglib_b200.models.DyGFormer import TransformerEncoder
    m = MergeLayer(272, 172, 172, 1).cuda().eval()
    sd = deterministic_state_dict(m.state_dict(), 6)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(333, 272, generator=g), torch.randn(333, 172, generator=g)
    close(m(a.cuda(), b.cuda()), om.merge_layer(sd, '', a, b), rtol=1e-4, atol=1e-5)
    tr = TransformerEncoder(200, 2, 0.1).cuda().eval()
    sd = deterministic_state_dict(tr.state_dict(), 7)
    tr.load_state_dict(sd)
    x = torch.randn(21, 64, 200, generator=g)
    o = om.OracleDyGFormer({'transformers.0.' + k: v for k, v in sd.items()}, np.zeros((2, 4), np.float32),
                           np.zeros((2, 4), np.float32), None, 50, 1, 1, 2, 8)
    close(tr(x.cuda()), o.transformer(0, x), rtol=1e-4, atol=2e-5)
    x = torch.randn(5, 17, 200, generator=g)                # ragged token count
    close(tr(x.cuda()), o.transformer(0, x), rtol=1e-4, atol=2e-5)


def test_cooc_encoder_forward():
    from dyglib_b200.models.DyGFormer import NeighborCooccurrenceEncoder
    enc = NeighborCooccurrenceEncoder(50, 'cuda').cuda().eval()
    sd = deterministic_state_dict(enc.state_dict(), 8)
    enc.load_state_dict(sd)
    rng = np.random.default_rng(0)
    s, d = rng.integers(0, 9, (23, 32)), rng.integers(0, 9, (23, 16))
    o = om.OracleDyGFormer({'neighbor_co_occurrence_encoder.' + k: v for k, v in sd.items()}, np.zeros((2, 4), np.float32),
                           np.zeros((2, 4), np.float32), None, 50, 1, 1, 2, 8)
    ws, wd = o.cooc_features(s, d)
    gs, gd = enc(s, d)
    close(gs, ws, rtol=1e-5, atol=1e-6)
    close(gd, wd, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('which', ['tgat', 'dygformer', 'TGN', 'DyRep', 'JODIE'])
def test_models_match_golden_and_oracle(which):
    _, tgat, dygformer, memory = cuda_factories()
    got = run_model_cases(tgat, dygformer, memory, which=(which,))
    gold = load_golden('models.npz')
    assert len(got) > 0
    worst = 0.0
    for k in got:
        close(got[k], gold[k], k)
        worst = max(worst, float(np.abs(got[k] - gold[k]).max()))
    _, otgat, odyg, omem = oracle_factories()
    want = run_model_cases(otgat, odyg, omem, which=(which,))
    for k in got:
        close(got[k], want[k], k)
    print(f'{which}: max abs diff vs reference golden = {worst:.3e}')


def test_dygformer_getters_match_oracle():
    sampler, _, dygformer, _ = cuda_factories()
    osampler, _, odyg, _ = oracle_factories()
    g = small_graph(seed=12)
    m, o = dygformer(g, 2, 16, 2), odyg(g, 2, 16, 2)
    src, dst, t, _, _ = next(batches(g, 1500, 1, 50))
    pn, pe, pt = m.pad_sequences(src, t, patch_size=2, max_input_sequence_length=16)
    wn, we, wt = o.padded(src, t)
    assert np.array_equal(pn, wn) and np.array_equal(pe, we) and np.array_equal(pt, wt)
    nf, ef, tf = m.get_features(t, pn, pe, pt, m.time_encoder)
    wnf, wef, wtf = o.features(t, wn, we, wt)
    assert np.array_equal(nf.cpu().numpy(), wnf.numpy()) and np.array_equal(ef.cpu().numpy(), wef.numpy())
    close(tf, wtf, rtol=0, atol=5e-7)


def test_tgn_many_batches_vs_oracle():
    """>= 50 consecutive batches so that memory drift would show."""
    _, _, _, memory = cuda_factories()
    _, _, _, omemory = oracle_factories()
    g = small_graph(seed=17, E=4000, nu=80, ni=30)
    (m, mem_fn), (o, omem_fn) = memory(g, 'TGN', 4), omemory(g, 'TGN', 4)
    with torch.no_grad():
        for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 60, 50)):
            a = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            b = o.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            c = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            d = o.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            if bi % 10 == 9:
                for x, y in zip(a + c, b + d):
                    close(x, y, f'batch {bi}')
    close(mem_fn(m)[0], omem_fn(o)[0], 'memory')
    assert np.array_equal(mem_fn(m)[1].cpu().numpy(), omem_fn(o)[1].numpy())
    m.assert_time_order()
    # exported pending messages have the reference's dict-of-lists form and the oracle's content
    exported = m.memory_bank.node_raw_messages
    pend = {int(v): lst for v, lst in o.raw_messages.items() if len(lst) > 0}
    assert set(exported) == set(pend)
    v = next(iter(pend))
    close(exported[v][0][0], pend[v][-1][0], 'message')
    assert float(exported[v][0][1]) == float(pend[v][-1][1])


@pytest.mark.parametrize('name', ['TGN', 'DyRep', 'JODIE'])
def test_memory_model_fused_pos_neg_call_equals_two_calls(name):
    """compute_pos_neg_temporal_embeddings (one embedding pass over the 4 B roots) == the reference loop's negative call followed
    by its positive call, batch after batch, including the memory it leaves behind."""
    _, _, _, memory = cuda_factories()
    g = small_graph(seed=19, E=3000, nu=70, ni=25)
    (m1, mem_fn), (m2, _) = memory(g, name, 4), memory(g, name, 4)
    with torch.no_grad():
        for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 30, 40)):
            a, b = m1.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            c, d = m1.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            fa, fb, fc, fd = m2.compute_pos_neg_temporal_embeddings(src, dst, neg, t, eid, 10)
            for x, y in zip((a, b, c, d), (fa, fb, fc, fd)):
                if name == 'TGN':    # one-launch step (dyg_tgn_step, BF16x3 mma tiles) vs kernel-per-step negative call (fp32 FFMA at B = 40)
                    close(x, y, f'{name} {bi}', rtol=1e-3, atol=5e-5)
                else:
                    assert torch.equal(x, y), (name, bi)
    if name == 'TGN':
        close(mem_fn(m1)[0], mem_fn(m2)[0], rtol=1e-3, atol=5e-5)
    else:
        assert torch.equal(mem_fn(m1)[0], mem_fn(m2)[0])
    assert torch.equal(mem_fn(m1)[1], mem_fn(m2)[1])


@pytest.mark.parametrize('B,k', [(40, 10), (200, 10), (7, 3), (33, 20)])
def test_tgn_one_launch_step_equals_kernel_per_step_path(B, k):
    """dyg_tgn_step (one cooperative launch per batch: search, attention, LayerNorm, MergeLayer, link predictor, persist,
    last-message election, messages, GRU, commit) against the same batch through the separate kernels, 25 consecutive batches:
    embeddings, link probabilities, memories, last-update times, pending messages."""
    from dyglib_b200.models.modules import MergeLayer
    _, _, _, memory = cuda_factories()
    g = small_graph(seed=19, E=6000, nu=70, ni=25)
    (m1, mem_fn), (m2, _) = memory(g, 'TGN', 4), memory(g, 'TGN', 4)
    m2.fused_step = False
    pred = MergeLayer(172, 172, 172, 1).cuda().eval()
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    assert m1._fused_step_ok(B)
    with torch.no_grad():
        for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 25, B)):
            f = m1.compute_pos_neg_temporal_embeddings(src, dst, neg, t, eid, k, link_predictor=pred)
            u = m2.compute_pos_neg_temporal_embeddings(src, dst, neg, t, eid, k, link_predictor=pred)
            assert len(f) == len(u) == 6
            for x, y in zip(f, u):
                close(x, y, f'batch {bi}', rtol=1e-3, atol=5e-5)     # fp32 FFMA tiles vs BF16x3 tcgen05 GEMMs at B >= 128
    m1.assert_time_order()
    close(mem_fn(m1)[0], mem_fn(m2)[0], rtol=1e-3, atol=5e-5)
    assert torch.equal(mem_fn(m1)[1], mem_fn(m2)[1])
    s1, s2 = m1.memory_bank._ensure(), m2.memory_bank._ensure()
    assert torch.equal(s1['pending'], s2['pending']) and torch.equal(s1['msg_time'], s2['msg_time']) and torch.equal(s1['lu_view'], s2['lu_view'])
    assert int(s1['winner'].max().item()) == -1                                  # the election table is left clean
    assert all(int(sc['barrier'].abs().sum().item()) == 0 for sc in m1._step_scratch.values())   # and the grid barrier re-armed
    pend = s1['pending'].bool()
    close(s1['msg_store'][pend], s2['msg_store'][pend], rtol=1e-3, atol=5e-5)
    close(s1['mem_view'], s2['mem_view'], rtol=1e-3, atol=5e-5)
    # the positive call alone also runs as one launch
    src, dst, t, eid, neg = next(batches(g, 25 * B, 1, B))
    with torch.no_grad():
        a = m1.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, k)
        b = m2.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, k)
    for x, y in zip(a, b):
        close(x, y, rtol=1e-3, atol=5e-5)


def test_tgn_same_node_src_and_dst_in_batch():
    """Non-bipartite batch: the kept message is the last *appended* (src-role appends, then dst-role), which
    need not be the chronologically last."""
    _, _, _, memory = cuda_factories()
    _, _, _, omemory = oracle_factories()
    g = small_graph(seed=19, E=600, nu=12, ni=6)
    rng = np.random.default_rng(0)
    g.dst_node_ids = rng.integers(1, g.num_nodes, g.num_interactions)   # make it non-bipartite with self loops
    (m, mem_fn), (o, omem_fn) = memory(g, 'TGN', 9), omemory(g, 'TGN', 9)
    with torch.no_grad():
        for src, dst, t, eid, neg in batches(g, 0, 10, 40):
            c = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 5)
            d = o.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 5)
            for x, y in zip(c, d):
                close(x, y)
    close(mem_fn(m)[0], omem_fn(o)[0], 'memory')
    assert np.array_equal(mem_fn(m)[1].cpu().numpy(), omem_fn(o)[1].numpy())


def test_dygformer_full_size_batch_vs_oracle():
    """One reference-sized batch (200 events) at the BASELINE config (wiki-shaped graph, P=2, L=64)."""
    _, _, dygformer, _ = cuda_factories()
    _, _, odyg, _ = oracle_factories()
    g = make_config_graph('dygformer_wiki')
    m, o = dygformer(g, 2, 64, 2), odyg(g, 2, 64, 2)
    src, dst, t, _, neg = next(batches(g, 140_000, 1, 200))
    with torch.no_grad():
        for d in (dst, neg):
            a = m.compute_src_dst_node_temporal_embeddings(src, d, t)
            b = o.compute_src_dst_node_temporal_embeddings(src, d, t)
            for x, y in zip(a, b):
                close(x, y)


@pytest.mark.parametrize('attn,ln', [(0, True), (1, False), (1, True), (2, True)])
def test_dygformer_attention_paths_agree_with_oracle(monkeypatch, attn, ln):
    """Every implementation of the attention half of the block (mma.sync path, LayerNorm planes + GEMM + tcgen05 attention, LayerNorm
    fused into the GEMM, the one-kernel form) against the oracle on batches with full and short padded lengths."""
    import dyglib_b200.models.DyGFormer as mod
    monkeypatch.setattr(mod, 'FUSED_ATTN', attn)
    monkeypatch.setattr(mod, 'FUSED_LN', ln)
    _, _, dygformer, _ = cuda_factories()
    _, _, odyg, _ = oracle_factories()
    g = small_graph(seed=21, E=6000, nu=150, ni=40)
    m, o = dygformer(g, 2, 32, 2), odyg(g, 2, 32, 2)
    with torch.no_grad():
        for start, n in ((40, 30), (4000, 120)):
            src, dst, t, _, _ = next(batches(g, start, 1, n))
            a = m.compute_src_dst_node_temporal_embeddings(src, dst, t)
            b = o.compute_src_dst_node_temporal_embeddings(src, dst, t)
            for x, y in zip(a, b):
                close(x, y)


def test_dygformer_grouped_equals_per_batch():
    """batch_size= groups keep the reference's per-batch padding unit: identical to one call per batch."""
    _, _, dygformer, _ = cuda_factories()
    g = small_graph(seed=12, E=6000, nu=200, ni=40)
    m = dygformer(g, 2, 16, 2)
    bs, nb, start = 25, 8, 50          # early events -> batches with different padded lengths
    sl = slice(start, start + bs * nb)
    src, dst, t = g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl]
    with torch.no_grad():
        gs, gd = m.compute_src_dst_node_temporal_embeddings(src, dst, t, batch_size=bs)
        for b in range(nb):
            q = slice(b * bs, (b + 1) * bs)
            es, ed = m.compute_src_dst_node_temporal_embeddings(src[q], dst[q], t[q])
            assert torch.equal(gs[q], es) and torch.equal(gd[q], ed), b


def test_graphmixer_matches_golden_and_oracle():
    """GraphMixer on the device sampler (k = 20 recent neighbours + a 300-deep window for the node encoder), the fused gather-GEMM
    projection and the tcgen05 GEMMs: eval embeddings against the reference's golden vectors and the oracle."""
    from helpers import cuda_graphmixer, oracle_graphmixer, run_graphmixer_cases
    gold = load_golden('graphmixer.npz')
    got = run_graphmixer_cases(cuda_graphmixer())
    want = run_graphmixer_cases(oracle_graphmixer())
    for k in got:
        np.testing.assert_allclose(got[k], gold[k], rtol=1e-3, atol=2e-4, err_msg=k)
        np.testing.assert_allclose(got[k], want[k], rtol=1e-3, atol=2e-4, err_msg=k)


def test_graphmixer_training_step_matches_golden():
    from helpers import cuda_graphmixer_train_step, assert_grads_close
    gold = {k[len('train.'):]: v for k, v in load_golden('graphmixer.npz').items() if k.startswith('train.')}
    got = cuda_graphmixer_train_step()
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    assert_grads_close(got, gold, rtol=2e-3)


def test_tcl_matches_golden_and_oracle():
    """TCL on the device sampler, the time-encode kernels and the tcgen05 GEMMs: eval embeddings against the reference's golden
    vectors and the oracle."""
    from helpers import cuda_tcl, oracle_tcl, run_tcl_cases
    gold = load_golden('tcl.npz')
    got = run_tcl_cases(cuda_tcl())
    want = run_tcl_cases(oracle_tcl())
    for k in got:
        np.testing.assert_allclose(got[k], gold[k], rtol=1e-3, atol=2e-4, err_msg=k)
        np.testing.assert_allclose(got[k], want[k], rtol=1e-3, atol=2e-4, err_msg=k)


def test_tcl_training_step_matches_golden():
    from helpers import cuda_tcl_train_step, assert_grads_close
    gold = {k[len('train.'):]: v for k, v in load_golden('tcl.npz').items() if k.startswith('train.')}
    got = cuda_tcl_train_step()
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    # 5e-3 of each gradient's largest magnitude: TCL's blocks gate with ReLU, and a pre-activation within the forward pass's
    # ~1e-5 rounding difference of zero flips its gate (observed 2.5e-3 on transformers.1.linear_layers.0.bias, < 2e-3 elsewhere)
    assert_grads_close(got, gold, rtol=5e-3)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs at their full size (the graphs, batch size and hyper-parameters bench.py runs), CUDA vs the oracle.
def test_tgat_myket_full_size_batches_vs_oracle():
    """Config 1: TGAT, 2 layers, 20 recent neighbours, batch 200 on the 694 k-event myket-shaped graph; two batches from the last
    15 % of the stream, positive and negative pairs."""
    _, tgat, _, _ = cuda_factories()
    _, otgat, _, _ = oracle_factories()
    g = make_config_graph('tgat_myket')
    m, o = tgat(g, 1), otgat(g, 1)
    start = int(g.num_interactions * 0.85)
    with torch.no_grad():
        for src, dst, t, _, neg in batches(g, start, 2, 200):
            for d in (dst, neg):
                a = m.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                b = o.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                for x, y in zip(a, b):
                    close(x, y)


def test_tgn_reddit_full_size_60_batches_vs_oracle():
    """Config 3: TGN (1 layer, 10 recent neighbours, last-message aggregation + GRU) over 60 consecutive 200-event batches of the
    672 k-event reddit-shaped graph from zeroed memory, through the call bench.py times (negative + positive roots in one pass):
    embeddings of every 10th batch, then the final memory, last-update times and the set of nodes holding a pending message."""
    _, _, _, memory = cuda_factories()
    _, _, _, omemory = oracle_factories()
    g = make_config_graph('tgn_reddit')
    (m, mem_fn), (o, omem_fn) = memory(g, 'TGN', 3), omemory(g, 'TGN', 3)
    with torch.no_grad():
        for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 60, 200)):
            got = m.compute_pos_neg_temporal_embeddings(src, dst, neg, t, eid, 10)
            wa = o.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
            wb = o.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            if bi % 10 == 9 or bi >= 57:
                for x, y in zip(got, (wa[0], wa[1], wb[0], wb[1])):
                    close(x, y, f'batch {bi}')
    m.assert_time_order()
    close(mem_fn(m)[0], omem_fn(o)[0], 'memory')
    assert np.array_equal(mem_fn(m)[1].cpu().numpy(), omem_fn(o)[1].numpy())
    pend = sorted(int(v) for v, lst in o.raw_messages.items() if len(lst) > 0)
    assert len(pend) > 1000                                   # a realistic pending set
    assert sorted(m.memory_bank.node_raw_messages) == pend


def test_dygformer_lastfm_full_size_batch_vs_oracle():
    """Config 4: DyGFormer patch 16, sequence length 512 on the 1.29 M-event lastfm-shaped graph (times up to 1.37e8 > 2^24, so the
    float32 rounding of the neighbour times matters); one 200-event batch from the last 15 %, positive and negative pairs."""
    _, _, dygformer, _ = cuda_factories()
    _, _, odyg, _ = oracle_factories()
    g = make_config_graph('dygformer_lastfm')
    m, o = dygformer(g, 16, 512, 2), odyg(g, 16, 512, 2)
    src, dst, t, _, neg = next(batches(g, int(g.num_interactions * 0.9), 1, 200))
    with torch.no_grad():
        for d in (dst, neg):
            a = m.compute_src_dst_node_temporal_embeddings(src, d, t)
            b = o.compute_src_dst_node_temporal_embeddings(src, d, t)
            for x, y in zip(a, b):
                close(x, y)
