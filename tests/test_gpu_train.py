"""GPU: the training path (autograd through the fused kernels) against the reference's gradients (golden fixture) and the
oracle's, plus a direct check of dyg_temporal_attend_bwd with dropout multipliers against a float64 torch formulation."""
import numpy as np
import pytest
import torch

from helpers import (cuda_tgat_train_step, oracle_tgat_train_step, assert_grads_close, load_golden, cuda_memory_train_step,
                     oracle_memory_train_step, cuda_dygformer_train_step, oracle_dygformer_train_step, DYG_TRAIN_CASES)

pytestmark = pytest.mark.gpu


def test_tgat_training_step_matches_reference_golden():
    got = cuda_tgat_train_step(dropout=0.0)
    gold = load_golden('tgat_train.npz')
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    np.testing.assert_allclose(got['pos'], gold['pos'], rtol=1e-3, atol=2e-4)
    assert_grads_close(got, gold, rtol=2e-3)


def test_tgat_training_step_matches_oracle():
    assert_grads_close(cuda_tgat_train_step(dropout=0.0), oracle_tgat_train_step(), rtol=2e-3)


@pytest.mark.parametrize('name', ['TGN', 'DyRep', 'JODIE'])
def test_memory_model_training_step_matches_reference_golden_and_oracle(name):
    """8 batches advance the memory on the eval kernels (no gradients), then one training batch: the recomputed
    get_updated_memories, the embedding module and the link predictor give the reference's loss and parameter gradients."""
    got = cuda_memory_train_step(name)
    gold = {k[len(name) + 1:]: v for k, v in load_golden('memory_train.npz').items() if k.startswith(name + '.')}
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    np.testing.assert_allclose(got['pos'], gold['pos'], rtol=1e-3, atol=2e-4)
    assert_grads_close(got, gold, rtol=2e-3)
    assert_grads_close(got, oracle_memory_train_step(name), rtol=2e-3)


@pytest.mark.parametrize('P,L', DYG_TRAIN_CASES)
def test_dygformer_training_step_matches_reference_golden_and_oracle(P, L):
    got = cuda_dygformer_train_step(P, L)
    gold = {k[len(f'P{P}_L{L}.'):]: v for k, v in load_golden('dygformer_train.npz').items() if k.startswith(f'P{P}_L{L}.')}
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    np.testing.assert_allclose(got['pos'], gold['pos'], rtol=1e-3, atol=2e-4)
    assert_grads_close(got, gold, rtol=2e-3)
    assert_grads_close(got, oracle_dygformer_train_step(P, L), rtol=2e-3)


def test_time_encode_backward_against_float64():
    """dyg_time_encode_bwd: gradients of w / b against autograd through cos of the same fp32-rounded argument in float64."""
    from dyglib_b200 import autograd as ag
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(3)
    n, T = 3001, 100
    dt = torch.rand(n, device=dev, generator=g) * 2e6
    w = (1.0 / 10 ** torch.linspace(0, 9, T, device=dev)).reshape(T, 1).requires_grad_(True)
    b = (0.1 * torch.randn(T, device=dev, generator=g)).requires_grad_(True)
    go = torch.randn(n, T, device=dev, generator=g)
    out = ag.time_encode(dt, w, b)
    out.backward(go)
    w2, b2 = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    val = dt.double().unsqueeze(1) * w2.reshape(1, T) + b2
    arg = val + (torch.addcmul(b.detach(), dt.unsqueeze(1), w.detach().reshape(1, T)).double() - val).detach()
    want = torch.cos(arg)
    # the comparison argument is rounded to fp32 like the kernel's FMA (torch's CUDA addcmul may or may not fuse: keep rows that agree)
    ok = (out.detach().double() - want.detach()).abs().max(dim=1).values < 1e-5
    assert float(ok.float().mean()) > 0.9
    (want * go.double() * ok.unsqueeze(1)).sum().backward()
    out2 = ag.time_encode(dt, w.detach().requires_grad_(True), b.detach().requires_grad_(True))
    ws, bs = out2.grad_fn.next_functions[1][0].variable, out2.grad_fn.next_functions[2][0].variable
    out2.backward(go * ok.unsqueeze(1))
    for got_, want_ in ((ws.grad.reshape(-1), w2.grad.reshape(-1)), (bs.grad, b2.grad)):
        scale = float(want_.abs().max())
        assert float((got_.double() - want_).abs().max()) / scale < 1e-3


def test_memory_model_eval_after_training_uses_current_weights():
    """A training call marks the look-ahead view stale; the next eval call rebuilds it, so eval after a weight change equals a
    fresh model with the new weights that replays the same stream."""
    from helpers import cuda_factories, small_graph, batches
    _, _, _, memory = cuda_factories()
    g = small_graph(seed=13)
    bs = list(batches(g, 0, 6, 30))

    def run(m, train_last):
        with torch.no_grad():
            for src, dst, t, eid, neg in bs[:4]:
                m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
        src, dst, t, eid, neg = bs[4]
        if train_last:
            m.train()
            m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
            m.eval()
        else:
            with torch.no_grad():
                m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
        with torch.no_grad():
            cell = m.memory_updater.memory_updater
            cell.weight_ih.mul_(1.5)
            # no manual invalidation: the bank's view key names the cell weights (their version counters moved)
            src, dst, t, eid, neg = bs[5]
            a, b = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
        return a.cpu().numpy(), b.cpu().numpy()
    m1, _ = memory(g, 'TGN', 3)
    m2, _ = memory(g, 'TGN', 3)
    for mod in list(m1.modules()) + list(m2.modules()):
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    a1, b1 = run(m1, True)
    a2, b2 = run(m2, False)
    np.testing.assert_allclose(a1, a2, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(b1, b2, rtol=1e-4, atol=1e-5)


def test_tgat_training_step_with_dropout_runs_and_differs():
    torch.manual_seed(0)
    a = cuda_tgat_train_step(dropout=0.1)
    torch.manual_seed(0)
    b = cuda_tgat_train_step(dropout=0.1)
    c = cuda_tgat_train_step(dropout=0.0)
    assert np.isfinite(a['loss']) and all(np.isfinite(v).all() for v in a.values())
    np.testing.assert_allclose(a['loss'], b['loss'], rtol=1e-6)       # same seed, same masks
    assert abs(float(a['loss']) - float(c['loss'])) > 1e-6             # dropout changes the step


@pytest.mark.parametrize('H,dense,drop', [(2, False, False), (2, True, True), (1, True, False), (2, False, True)])
def test_temporal_attend_backward_against_float64(H, dense, drop):
    """s_h = sum_j softmax_j(q_h . x_j) m_hj x_j with x_j = [node | edge | cos(dt w + b)]: gradients of q, dense node rows,
    w and b from the kernel vs autograd through the same formula in float64."""
    from dyglib_b200 import autograd as ag
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(7)
    n, k, F, E, T = 37, 20, 172, 172, 100
    Dk = F + E + T
    node_tab = torch.randn(50, F, device=dev, generator=g) * 0.3
    edge_tab = torch.randn(80, E, device=dev, generator=g) * 0.3
    node_tab[0] = 0
    edge_tab[0] = 0
    nidx = torch.randint(0, 50, (n * k,), device=dev, generator=g)
    nidx[:k] = 0                                                           # a fully masked root
    eidx = torch.randint(0, 80, (n * k,), device=dev, generator=g)
    eidx[nidx == 0] = 0
    tq = torch.rand(n, device=dev, generator=g, dtype=torch.float64) * 1e5 + 1e5
    tn = (tq.float().repeat_interleave(k) - torch.rand(n * k, device=dev, generator=g) * 1e5).contiguous()
    w = (1.0 / 10 ** torch.linspace(0, 9, T, device=dev)).reshape(T, 1).requires_grad_(True)
    b = (0.1 * torch.randn(T, device=dev, generator=g)).requires_grad_(True)
    qk = (0.05 * torch.randn(n, H * Dk, device=dev, generator=g)).requires_grad_(True)
    nbr_dense = (torch.randn(n * k, F, device=dev, generator=g) * 0.3).requires_grad_(True) if dense else None
    gs = torch.randn(n, H * Dk, device=dev, generator=g)
    if drop:
        torch.manual_seed(3)
    s = ag.temporal_attend(qk, nbr_dense, w, b, n=n, k=k, H=H, node_tab=node_tab, node_idx=nidx, F=F, edge_tab=edge_tab,
                           edge_idx=eidx, E=E, T=T, mask_ids=nidx, t_query=tq, t_nbr=tn, zero_row0=3, dropout=0.25 if drop else 0.0)
    s.backward(gs)
    got = [qk.grad.clone(), w.grad.clone(), b.grad.clone()] + ([nbr_dense.grad.clone()] if dense else [])
    # float64 formulation
    if drop:
        torch.manual_seed(3)
        m = ((torch.rand((n, H, k), device=dev) >= 0.25).float() / 0.75).double()
    else:
        m = torch.ones((n, H, k), device=dev, dtype=torch.float64)
    qk2, w2, b2 = qk.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    nd2 = nbr_dense.detach().double().requires_grad_(True) if dense else None
    nodes = nd2 if dense else node_tab.double()[nidx]
    dt = (tq.repeat_interleave(k) - tn.double()).float().double()
    val = dt.unsqueeze(1) * w2.reshape(1, T) + b2
    arg = val + (val.float().double() - val).detach()          # the fp32-rounded argument of the forward pass, d arg = d val
    x = torch.cat([nodes, edge_tab.double()[eidx], torch.cos(arg)], dim=1).reshape(n, k, Dk)
    q = qk2.reshape(n, H, Dk)
    sc = torch.einsum('nhd,nkd->nhk', q, x).masked_fill((nidx == 0).reshape(n, 1, k), -1e10)
    a = torch.softmax(sc, dim=-1) * m
    want_s = torch.einsum('nhk,nkd->nhd', a, x).reshape(n, H * Dk)
    want_s.backward(gs.double())
    want = [qk2.grad, w2.grad, b2.grad] + ([nd2.grad] if dense else [])
    assert float((s.detach().double() - want_s.detach()).abs().max()) < 5e-5
    for name, a_, b_ in zip(['qk', 'w', 'b', 'nbr'], got, want):
        scale = max(float(b_.abs().max()), 1e-6)
        err = float((a_.double() - b_).abs().max()) / scale
        assert err < 1e-3, (name, err)
