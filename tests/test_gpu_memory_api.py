"""GPU: the MemoryModel sub-API of SURVEY.md section 8(b) (reference ``models/MemoryModel.py:170-251, 389-407, 435-487, 534-545``)
and the checkpoint / backup semantics of ``utils/EarlyStopping.py:65-86``, ``evaluate_link_prediction.py:152-156`` and
``train_link_prediction.py:265-299`` over the device tables, against the oracle and against the fused path."""
import io
from collections import defaultdict

import numpy as np
import pytest
import torch

from helpers import small_graph, batches, cuda_factories, oracle_factories

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-3, atol=2e-4)


def _run(m, bs, k=10):
    outs = []
    with torch.no_grad():
        for src, dst, t, eid, neg in bs:
            a = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, k)
            b = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, k)
            outs.append([x.detach().cpu().numpy() for x in (a[0], a[1], b[0], b[1])])
    return outs


@pytest.mark.parametrize('name', ['TGN', 'DyRep', 'JODIE'])
def test_checkpoint_round_trip_then_matches_oracle(name):
    """EarlyStopping.save_checkpoint / load_checkpoint: ``torch.save(state_dict)`` + ``torch.save(node_raw_messages)``, then
    ``load_state_dict`` + ``memory_bank.node_raw_messages = torch.load(...)`` into a fresh model; the next 5 batches must match
    the oracle that ran all 15 batches (and the model that was never checkpointed)."""
    _, _, _, cmem = cuda_factories()
    _, _, _, omem = oracle_factories()
    g = small_graph(seed=13)
    bs = list(batches(g, 0, 15, 30))
    a, _ = cmem(g, name, 3)
    _run(a, bs[:10])
    f_sd, f_msg = io.BytesIO(), io.BytesIO()
    torch.save(a.state_dict(), f_sd)
    torch.save(a.memory_bank.node_raw_messages, f_msg)
    f_sd.seek(0)
    f_msg.seek(0)
    raw = torch.load(f_msg, map_location='cpu', weights_only=False)
    assert type(raw) is defaultdict and len(raw) > 0               # the reference's plain format
    v0 = next(iter(raw))
    assert isinstance(raw[v0], list) and raw[v0][0][0].shape == (a.message_dim,)
    b, _ = cmem(g, name, 4)                                         # other weights, empty memory
    b.load_state_dict(torch.load(f_sd, map_location='cpu'))
    b.memory_bank.node_raw_messages = raw
    # evaluate_link_prediction.py:152-156: move every stored message to the device, item by item
    for node_id, msgs in b.memory_bank.node_raw_messages.items():
        b.memory_bank.node_raw_messages[node_id] = [(mm[0].to('cuda'), mm[1]) for mm in msgs]
    assert len(b.memory_bank.node_raw_messages) == len(raw)
    got_b = _run(b, bs[10:])
    got_a = _run(a, bs[10:])
    o, _ = omem(g, name, 3)
    want = _run(o, bs)[10:]
    for ga, gb, w in zip(got_a, got_b, want):
        for x, y, z in zip(ga, gb, w):
            np.testing.assert_allclose(x, z, **TOL)
            np.testing.assert_allclose(y, z, **TOL)
    np.testing.assert_allclose(b.memory_bank.node_memories.data.cpu().numpy(), o.memory.numpy(), **TOL)
    np.testing.assert_array_equal(b.memory_bank.node_last_updated_times.data.cpu().numpy(), o.last_update.numpy())
    pend_o = sorted(int(v) for v, lst in o.raw_messages.items() if len(lst))
    assert sorted(b.memory_bank.node_raw_messages) == pend_o


def test_sub_api_flow_equals_fused_call():
    """The reference's positive-batch body driven through the sub-API (get_updated_memories, update_memories,
    clear_node_raw_messages, compute_new_node_raw_messages, store_node_raw_messages) leaves the same state as the fused call."""
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    bs = list(batches(g, 0, 8, 30))
    a, _ = cmem(g, 'TGN', 3)
    b, _ = cmem(g, 'TGN', 3)
    _run(a, bs)
    with torch.no_grad():
        for src, dst, t, eid, neg in bs:
            bank = b.memory_bank
            node_ids = np.concatenate([src, dst])
            mem_all, lu_all = b.get_updated_memories(np.arange(b.num_nodes), bank.node_raw_messages)
            # the same through the reference's host dict format
            mem_d, lu_d = b.get_updated_memories(np.arange(b.num_nodes), bank.node_raw_messages.to_dict())
            np.testing.assert_allclose(mem_all.cpu().numpy(), mem_d.cpu().numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_array_equal(lu_all.cpu().numpy(), lu_d.cpu().numpy())
            b.update_memories(node_ids, bank.node_raw_messages)
            bank.clear_node_raw_messages(node_ids)
            us, ms = b.compute_new_node_raw_messages(src, dst, None, t, eid)
            ud, md = b.compute_new_node_raw_messages(dst, src, None, t, eid)
            bank.store_node_raw_messages(us, ms)
            bank.store_node_raw_messages(ud, md)
    # a: one-launch step (BF16x3 mma tiles), b: fused fp32 cell through the sub-API -- the same state within the dense-layer tolerance
    for x, y in ((a.memory_bank.node_memories.data, b.memory_bank.node_memories.data),
                 (a.memory_bank.node_last_updated_times.data, b.memory_bank.node_last_updated_times.data)):
        np.testing.assert_allclose(x.cpu().numpy(), y.cpu().numpy(), rtol=1e-3, atol=5e-5)
    ra, rb = a.memory_bank.node_raw_messages.to_dict(), b.memory_bank.node_raw_messages.to_dict()
    assert sorted(ra) == sorted(rb)
    for v in ra:
        assert ra[v][-1][1] == rb[v][-1][1]
        np.testing.assert_allclose(ra[v][-1][0].cpu().numpy(), rb[v][-1][0].cpu().numpy(), rtol=1e-3, atol=5e-5)
    # and both continue identically
    nxt = list(batches(g, 240, 2, 30))
    for x, y in zip(_run(a, nxt), _run(b, nxt)):
        for p, q in zip(x, y):
            np.testing.assert_allclose(p, q, rtol=1e-3, atol=5e-5)


def test_memory_updater_compat_calls_match_torch_cells():
    """MemoryUpdater.get_updated_memories / update_memories on explicit (ids, messages, timestamps): one fused
    dyg_gru_update_fwd launch against torch's GRUCell / RNNCell on the CPU."""
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    rng = np.random.default_rng(0)
    for name in ('TGN', 'DyRep'):
        m, _ = cmem(g, name, 3)
        bank = m.memory_bank
        bank.node_memories.data.copy_(torch.from_numpy(rng.standard_normal(tuple(bank.node_memories.shape)).astype(np.float32)))
        ids = np.unique(rng.integers(1, m.num_nodes, 37))
        msgs = torch.from_numpy(rng.standard_normal((len(ids), m.message_dim)).astype(np.float32)).cuda()
        ts = np.arange(len(ids), dtype=np.float64) + 5.0
        cell = m.memory_updater.memory_updater
        ref_cell = type(cell)(m.message_dim, m.memory_dim)
        ref_cell.load_state_dict({k: v.cpu() for k, v in cell.state_dict().items()})
        h0 = bank.node_memories.data.cpu().clone()
        with torch.no_grad():
            want = ref_cell(msgs.cpu(), h0[torch.from_numpy(ids)])
            mem, lu = m.memory_updater.get_updated_memories(ids, msgs, ts)
        np.testing.assert_allclose(mem[torch.from_numpy(ids).cuda()].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
        rest = np.setdiff1d(np.arange(m.num_nodes), ids)
        np.testing.assert_array_equal(mem[torch.from_numpy(rest).cuda()].cpu().numpy(), h0[torch.from_numpy(rest)].numpy())
        np.testing.assert_array_equal(bank.node_memories.data.cpu().numpy(), h0.numpy())          # nothing persisted
        np.testing.assert_array_equal(lu[torch.from_numpy(ids).cuda()].cpu().numpy(), ts.astype(np.float32))
        with torch.no_grad():
            m.memory_updater.update_memories(ids, msgs, ts)
        np.testing.assert_allclose(bank.node_memories.data[torch.from_numpy(ids).cuda()].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_array_equal(bank.node_last_updated_times.data[torch.from_numpy(ids).cuda()].cpu().numpy(), ts.astype(np.float32))
        with pytest.raises(AssertionError):          # models/MemoryModel.py:448-449
            m.memory_updater.update_memories(ids, msgs, ts - 100.0)


def test_reload_of_backup_taken_under_other_weights_rebuilds_the_view():
    """train -> backup -> (weights move) -> reload must not reuse the backup's look-ahead view."""
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    bs = list(batches(g, 0, 6, 30))

    def run(with_backup):
        m, _ = cmem(g, 'TGN', 3)
        _run(m, bs[:4])
        bk = m.memory_bank.backup_memory_bank() if with_backup else None
        if with_backup:
            _run(m, bs[4:5])                         # e.g. a validation pass that advances the memory
        with torch.no_grad():
            m.memory_updater.memory_updater.weight_ih.mul_(1.25)
        if with_backup:
            m.memory_bank.reload_memory_bank(bk)
        return _run(m, bs[5:6])[0]
    for x, y in zip(run(True), run(False)):
        np.testing.assert_allclose(x, y, rtol=1e-4, atol=1e-5)


def test_load_state_dict_resyncs_the_view():
    """memories loaded into a model that has already run must be what the next eval call reads."""
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    bs = list(batches(g, 0, 6, 30))
    a, _ = cmem(g, 'TGN', 3)
    _run(a, bs[:3])
    sd = {k: v.clone() for k, v in a.state_dict().items()}
    raw = a.memory_bank.node_raw_messages.to_dict()
    want = _run(a, bs[3:4])[0]
    b, _ = cmem(g, 'TGN', 3)
    _run(b, bs[:5])                                  # b has run further: its view is of another state
    b.load_state_dict(sd)
    b.memory_bank.node_raw_messages = raw
    got = _run(b, bs[3:4])[0]
    for x, y in zip(got, want):       # b's view was rebuilt by the fp32 cell, a's was maintained by the BF16x3 step
        np.testing.assert_allclose(x, y, rtol=1e-3, atol=5e-5)


def test_message_list_mutations_write_through():
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    m, _ = cmem(g, 'TGN', 3)
    bank = m.memory_bank
    store = bank.node_raw_messages
    assert len(store) == 0 and store[5] == [] and 5 not in store
    row = torch.arange(m.message_dim, dtype=torch.float32)
    store[5].append((row, np.float64(12.0)))
    store[7] = [(row * 2, 3.0), (row * 3, 4.0)]      # only the last message of a list can ever be aggregated
    assert sorted(bank.node_raw_messages) == [5, 7]
    assert bank.node_raw_messages[7][-1][1] == 4.0
    np.testing.assert_array_equal(bank.node_raw_messages[7][-1][0].cpu().numpy(), (row * 3).numpy())
    ids, msgs, ts = m.message_aggregator.aggregate_messages(np.arange(m.num_nodes), bank.node_raw_messages)
    assert ids.tolist() == [5, 7] and ts.tolist() == [12.0, 4.0] and tuple(msgs.shape) == (2, m.message_dim)
    bank.clear_node_raw_messages(np.array([5]))
    assert sorted(bank.node_raw_messages) == [7]
    del bank.node_raw_messages[7]
    assert len(bank.node_raw_messages) == 0
    with pytest.raises(KeyError):
        bank.node_raw_messages[m.num_nodes + 3]


def test_time_projection_embedding_compat():
    _, _, _, cmem = cuda_factories()
    g = small_graph(seed=13)
    m, _ = cmem(g, 'JODIE', 3)
    rng = np.random.default_rng(1)
    mem = torch.from_numpy(rng.standard_normal((m.num_nodes, m.memory_dim)).astype(np.float32)).cuda()
    ids = rng.integers(0, m.num_nodes, 50)
    iv = torch.from_numpy(rng.standard_normal(50).astype(np.float32)).cuda()
    em = m.embedding_module
    with torch.no_grad():
        got = em.compute_node_temporal_embeddings(mem, ids, iv)
        want = mem[torch.from_numpy(ids).cuda()] * (1 + em.linear_layer(iv.unsqueeze(1)))
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-6)
