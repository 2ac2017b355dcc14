"""GPU: the pieces composed the way the reference's driver composes them (train_link_prediction.py:97-263): files -> splits ->
samplers -> model.train() -> negatives -> BCE -> Adam -> metrics, and a CUDA-graph-replayed training step followed by eval."""
import numpy as np
import pytest
import torch

from test_host_plumbing import _write_dataset

pytestmark = pytest.mark.gpu


def _setup(tmp_path):
    from dyglib_b200.utils.DataLoader import get_link_prediction_data
    _write_dataset(str(tmp_path), 'toy', E=6000, nu=150, ni=60, seed=4)
    return get_link_prediction_data('toy', 0.15, 0.15, root=str(tmp_path), verbose=False)


def test_tgat_training_loop_end_to_end(tmp_path):
    from dyglib_b200.utils.DataLoader import get_idx_data_loader
    from dyglib_b200.utils.utils import get_neighbor_sampler, NegativeEdgeSampler, set_random_seed
    from dyglib_b200.utils.metrics import get_link_prediction_metrics
    from dyglib_b200.models.TGAT import TGAT
    from dyglib_b200.models.modules import MergeLayer
    nf, ef, full, train, val, test, _, _ = _setup(tmp_path)
    set_random_seed(0)
    train_sampler, full_sampler = get_neighbor_sampler(train, 'recent'), get_neighbor_sampler(full, 'recent')
    model = TGAT(nf, ef, train_sampler, 100, 2, 2, 0.1, 'cuda')
    pred = MergeLayer(172, 172, 172, 1).to('cuda')
    opt = torch.optim.Adam(list(model.parameters()) + list(pred.parameters()), lr=1e-3)
    neg_sampler = NegativeEdgeSampler(train.src_node_ids, train.dst_node_ids, seed=0)
    epoch_loss = []
    for epoch in range(3):
        model.train(); pred.train()
        model.set_neighbor_sampler(train_sampler)
        losses = []
        for idx in get_idx_data_loader(list(range(train.num_interactions)), batch_size=200, shuffle=False):
            idx = idx.numpy()
            src, dst, t = train.src_node_ids[idx], train.dst_node_ids[idx], train.node_interact_times[idx]
            _, neg = neg_sampler.sample(len(idx))
            ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t, 20)
            ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t, 20)
            p = torch.cat([pred(ps, pd).squeeze(-1).sigmoid(), pred(ns, nd).squeeze(-1).sigmoid()])
            y = torch.cat([torch.ones(len(idx), device='cuda'), torch.zeros(len(idx), device='cuda')])
            loss = torch.nn.functional.binary_cross_entropy(p, y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss.item()))
        assert np.isfinite(losses).all()
        epoch_loss.append(float(np.mean(losses)))
    assert epoch_loss[-1] < epoch_loss[0], epoch_loss
    # evaluation on the validation split with the full-graph sampler (evaluate_models_utils.py:19-100)
    model.eval(); pred.eval()
    model.set_neighbor_sampler(full_sampler)
    val_neg = NegativeEdgeSampler(full.src_node_ids, full.dst_node_ids, seed=0)
    aps = []
    with torch.no_grad():
        for idx in get_idx_data_loader(list(range(val.num_interactions)), batch_size=200, shuffle=False):
            idx = idx.numpy()
            src, dst, t = val.src_node_ids[idx], val.dst_node_ids[idx], val.node_interact_times[idx]
            _, neg = val_neg.sample(len(idx))
            ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t, 20)
            ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t, 20)
            p = torch.cat([pred(ps, pd).squeeze(-1).sigmoid(), pred(ns, nd).squeeze(-1).sigmoid()])
            y = torch.cat([torch.ones(len(idx)), torch.zeros(len(idx))]).cuda()
            m = get_link_prediction_metrics(p, y)
            assert 0.0 < m['average_precision'] <= 1.0 and 0.0 <= m['roc_auc'] <= 1.0
            aps.append(m['average_precision'])
    assert len(aps) > 0      # no claim about quality: the toy graph is uniform random, there is nothing to learn beyond chance


def test_graph_replayed_training_then_eval_sees_the_new_weights(tmp_path):
    """A replayed training graph moves the parameters without bumping their version counters: eval afterwards must rebuild the
    folded weights / operand planes / cached constants (ops.WEIGHTS_EPOCH) and equal a fresh model with the trained state."""
    from dyglib_b200.utils.utils import get_neighbor_sampler, set_random_seed
    from dyglib_b200.utils.dist import GradBucket
    from dyglib_b200.utils.graph import GraphedStep
    from dyglib_b200.models.TGAT import TGAT
    from dyglib_b200.models.modules import MergeLayer
    nf, ef, full, train, val, test, _, _ = _setup(tmp_path)
    set_random_seed(1)
    sampler = get_neighbor_sampler(full, 'recent')
    model = TGAT(nf, ef, sampler, 100, 2, 2, 0.0, 'cuda')
    pred = MergeLayer(172, 172, 172, 1).to('cuda')
    B = 100
    sl = slice(3000, 3000 + B)
    src = torch.from_numpy(full.src_node_ids[sl]).cuda()
    dst = torch.from_numpy(full.dst_node_ids[sl]).cuda()
    t = torch.from_numpy(full.node_interact_times[sl]).cuda()
    neg = torch.from_numpy(np.random.RandomState(0).choice(np.unique(full.dst_node_ids), B)).cuda()
    model.eval()
    with torch.no_grad():
        before = model.compute_node_temporal_embeddings(src, t, 2, 20).clone()       # builds every cache with the initial weights
    model.train(); pred.train()
    bucket = GradBucket(list(model.parameters()) + list(pred.parameters()))
    opt = torch.optim.Adam(bucket.params, lr=1e-2, capturable=True)

    def step(src, dst, neg, t):
        bucket.zero()
        emb = model.compute_node_temporal_embeddings(torch.cat([src, dst, src, neg]), torch.cat([t, t, t, t]), 2, 20)
        p = torch.cat([pred(emb[:B], emb[B:2 * B]).squeeze(-1).sigmoid(), pred(emb[2 * B:3 * B], emb[3 * B:]).squeeze(-1).sigmoid()])
        y = torch.cat([torch.ones(B, device='cuda'), torch.zeros(B, device='cuda')])
        loss = torch.nn.functional.binary_cross_entropy(p, y)
        loss.backward()
        opt.step()
        return loss.detach()
    graphed = GraphedStep(step, (src, dst, neg, t), warmup=3, grad=True)
    losses = [float(graphed(src, dst, neg, t).item()) for _ in range(5)]
    assert losses[-1] < losses[0]
    bucket.check_views()
    model.eval()
    with torch.no_grad():
        after = model.compute_node_temporal_embeddings(src, t, 2, 20)
        fresh = TGAT(nf, ef, sampler, 100, 2, 2, 0.0, 'cuda').eval()
        fresh.load_state_dict(model.state_dict())
        want = fresh.compute_node_temporal_embeddings(src, t, 2, 20)
    assert not torch.allclose(before, after, atol=1e-4)
    assert torch.equal(after, want)


def test_graphed_step_input_paths_agree():
    """GraphedStep copies its inputs into one captured device slab: device tensors (one multi-tensor copy per dtype), pinned host
    tensors (packed into a pinned slab ring, ONE H2D copy) and pageable host tensors (per-tensor copies) must replay the same step,
    also when the ring of pinned slabs wraps around while earlier copies are still in flight."""
    from dyglib_b200.utils.graph import GraphedStep
    g = torch.Generator().manual_seed(0)
    n = 257
    mk = lambda: (torch.randint(0, 1000, (n,), generator=g), torch.randint(0, 1000, (n,), generator=g),   # noqa: E731
                  torch.rand(n, generator=g, dtype=torch.float64), torch.randint(0, 1000, (5, 3), generator=g))
    w = torch.arange(n, device='cuda', dtype=torch.float64)

    def step(a, b, t, c):
        return (a.double() * 3.0 + b.double()) * t + w + c.double().sum()
    first = tuple(x.cuda() for x in mk())
    graphed = GraphedStep(step, first, warmup=1)
    assert torch.equal(graphed(*first), step(*first))
    for i in range(20):                     # more calls than ring slots
        host = mk()
        want = step(*[x.cuda() for x in host])
        got_pinned = graphed(*[x.pin_memory() for x in host]).clone()
        got_pageable = graphed(*host).clone()
        got_device = graphed(*[x.cuda() for x in host]).clone()
        assert torch.equal(got_pinned, want) and torch.equal(got_pageable, want) and torch.equal(got_device, want), i
