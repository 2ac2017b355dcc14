"""Shared test fixtures: small seeded graphs, deterministic weights, batch iteration."""
import numpy as np
import torch

from dyglib_b200.synthetic import make_graph


def small_graph(seed=7, E=3000, nu=60, ni=25, tmax=200000.0, F=172):
    return make_graph(E, nu, ni, tmax, seed, feat_dim=F)


def deterministic_state_dict(template: dict, seed: int = 0) -> dict:
    """Weights that depend only on (key, shape, seed), so the reference (build container), the oracle and the
    CUDA modules (GPU box) can be given identical parameters without shipping checkpoints."""
    import zlib
    out = {}
    for key in sorted(template.keys()):
        shape = tuple(template[key].shape)
        # shared modules appear under two prefixes in the reference's state_dict; give aliases one value
        canon = key.replace('embedding_module.time_encoder', 'time_encoder').replace('memory_updater.memory_bank', 'memory_bank')
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(canon.encode()))
        if 'node_memories' in key or 'node_last_updated_times' in key:
            v = torch.zeros(shape)
        elif key.endswith('time_encoder.w.weight'):
            v = template[key].detach().clone().float().cpu()          # keep the fixed 1/10^linspace init
        elif key.endswith('time_encoder.w.bias'):
            v = 0.1 * torch.randn(shape, generator=g)                 # non-zero bias exercises the fp32 FMA
        elif ('norm' in key) and key.endswith('weight'):
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            v = torch.randn(shape, generator=g) / np.sqrt(shape[-1])
        else:
            v = 0.1 * torch.randn(shape, generator=g)
        out[key] = v.float()
    return out


def batches(g, start, nb, B, seed=0):
    rng = np.random.RandomState(seed)
    uniq = np.unique(g.dst_node_ids)
    for b in range(nb):
        sl = slice(start + b * B, start + (b + 1) * B)
        neg = uniq[rng.randint(0, len(uniq), B)]
        yield g.src_node_ids[sl], g.dst_node_ids[sl], g.node_interact_times[sl], g.edge_ids[sl], neg


def make_queries(g, n, rng, with_f32=False):
    e = rng.integers(0, g.num_interactions, n)
    side = rng.integers(0, 2, n).astype(bool)
    nodes = np.where(side, g.src_node_ids[e], g.dst_node_ids[e])
    times = g.node_interact_times[e].copy()
    nodes[: n // 20] = 0
    times[n // 20: n // 10] += 0.5
    if with_f32:
        times = times.astype(np.float32)
    return nodes, times


# ---------------------------------------------------------------------------------------------
# Case runners shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle / CUDA-vs-golden (GPU) tests.
# `make_sampler(graph, strategy, seed, tsf)` returns an object with the reference's sampler API.
def run_sampler_cases(make_sampler, pad_fn, cooc_fn):
    g = small_graph(seed=7)
    out = {}
    for strategy, seed in (('recent', None), ('uniform', 3), ('time_interval_aware', 3)):
        s = make_sampler(g, strategy, seed, 1e-5)
        rng = np.random.default_rng(0)
        for k, f32 in ((20, False), (3, True), (1, False)):
            nodes, times = make_queries(g, 400, rng, f32)
            a, b, c = s.get_historical_neighbors(nodes, times, k)
            out[f'{strategy}_k{k}_nbr'], out[f'{strategy}_k{k}_eid'], out[f'{strategy}_k{k}_t'] = a, b, c
    s = make_sampler(g, 'recent', None, 0.0)
    rng = np.random.default_rng(5)
    nodes, times = make_queries(g, 100, rng)
    ln, le, lt = s.get_multi_hop_neighbors(2, nodes, times, 3)
    for h in range(2):
        out[f'multihop_{h}_nbr'], out[f'multihop_{h}_eid'], out[f'multihop_{h}_t'] = ln[h], le[h], lt[h]
    rng = np.random.default_rng(1)
    n1, t1 = make_queries(g, 200, rng)
    n2, _ = make_queries(g, 200, rng)
    pads = []
    for nodes in (n1, n2):
        a = s.get_all_first_hop_neighbors(nodes, t1)
        out.setdefault('firsthop_len', np.array([len(x) for x in a[0]]))
        pads.append(pad_fn(s, g, nodes, t1, a, 4, 32))
    for i, p in enumerate(pads):
        out[f'pad{i}_nbr'], out[f'pad{i}_eid'], out[f'pad{i}_t'] = p
    out['cooc_src'], out['cooc_dst'] = cooc_fn(pads[0][0], pads[1][0])
    return out


def run_model_cases(make_tgat, make_dygformer, make_memory, which=('tgat', 'dygformer', 'TGN', 'DyRep', 'JODIE')):
    """make_*(graph, ...) return objects exposing compute_src_dst_node_temporal_embeddings; outputs as numpy."""
    out = {}

    def np_(x):
        return x.detach().cpu().numpy()
    with torch.no_grad():
        if 'tgat' in which:
            g = small_graph(seed=11)
            m = make_tgat(g, 1)
            for bi, (src, dst, t, _, neg) in enumerate(batches(g, 2000, 2, 40)):
                for tag, d in (('pos', dst), ('neg', neg)):
                    a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t, 20)
                    out[f'tgat_{bi}_{tag}_src'], out[f'tgat_{bi}_{tag}_dst'] = np_(a), np_(b)
        if 'dygformer' in which:
            g = small_graph(seed=12)
            for P, L in ((2, 16), (1, 8), (4, 32)):
                m = make_dygformer(g, P, L, 2)
                for bi, (src, dst, t, _, neg) in enumerate(batches(g, 1000, 2, 50)):
                    for tag, d in (('pos', dst), ('neg', neg)):
                        a, b = m.compute_src_dst_node_temporal_embeddings(src, d, t)
                        out[f'dygformer_P{P}_L{L}_{bi}_{tag}_src'], out[f'dygformer_P{P}_L{L}_{bi}_{tag}_dst'] = np_(a), np_(b)
        g = small_graph(seed=13)
        for name in ('TGN', 'DyRep', 'JODIE'):
            if name not in which:
                continue
            m, mem_fn = make_memory(g, name, 3)
            for bi, (src, dst, t, eid, neg) in enumerate(batches(g, 0, 12, 30)):
                ra = m.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, 10)
                rb = m.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, 10)
                if bi >= 9:
                    out[f'{name}_{bi}_neg_src'], out[f'{name}_{bi}_neg_dst'] = np_(ra[0]), np_(ra[1])
                    out[f'{name}_{bi}_pos_src'], out[f'{name}_{bi}_pos_dst'] = np_(rb[0]), np_(rb[1])
            mem, lu = mem_fn(m)
            out[f'{name}_memory'], out[f'{name}_last_update'] = np_(mem), np_(lu)
    return out


def templates():
    """state_dict templates (keys / shapes) from the CUDA package's parameter containers, built on CPU."""
    from dyglib_b200.models.TGAT import TGAT
    from dyglib_b200.models.DyGFormer import DyGFormer
    from dyglib_b200.models.MemoryModel import MemoryModel
    return TGAT, DyGFormer, MemoryModel


def oracle_factories():
    from oracle.sampler import OracleSampler
    from oracle.models import OracleTGAT, OracleDyGFormer, OracleMemoryModel
    TGAT, DyGFormer, MemoryModel = templates()

    def sampler(g, strategy='recent', seed=None, tsf=0.0):
        return OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, strategy, tsf, seed)

    def tgat(g, wseed):
        sd = deterministic_state_dict(TGAT(g.node_raw_features, g.edge_raw_features, None, 100, 2, 2, 0.1, 'cpu').state_dict(), wseed)
        return OracleTGAT(sd, g.node_raw_features, g.edge_raw_features, sampler(g), 2, 2)

    def dygformer(g, P, L, wseed):
        sd = deterministic_state_dict(DyGFormer(g.node_raw_features, g.edge_raw_features, None, 100, 50, P, 2, 2, 0.1, L, 'cpu').state_dict(), wseed)
        return OracleDyGFormer(sd, g.node_raw_features, g.edge_raw_features, sampler(g), 50, P, 2, 2, L)

    def memory(g, name, wseed):
        sd = deterministic_state_dict(MemoryModel(g.node_raw_features, g.edge_raw_features, None, 100, name, 1, 2, 0.1, device='cpu').state_dict(), wseed)
        m = OracleMemoryModel(sd, g.node_raw_features, g.edge_raw_features, sampler(g), name, 1, 2, 3.0, 50.0, 5.0, 70.0)
        return m, (lambda mm: (mm.memory, mm.last_update))
    return sampler, tgat, dygformer, memory


def cuda_factories():
    from dyglib_b200.utils.utils import get_neighbor_sampler
    TGAT, DyGFormer, MemoryModel = templates()

    def sampler(g, strategy='recent', seed=None, tsf=0.0):
        return get_neighbor_sampler(g, strategy, time_scaling_factor=tsf, seed=seed)

    def tgat(g, wseed):
        m = TGAT(g.node_raw_features, g.edge_raw_features, sampler(g), 100, 2, 2, 0.1, 'cuda').eval()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), wseed))
        return m

    def dygformer(g, P, L, wseed):
        m = DyGFormer(g.node_raw_features, g.edge_raw_features, sampler(g), 100, 50, P, 2, 2, 0.1, L, 'cuda').eval()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), wseed))
        return m

    def memory(g, name, wseed):
        m = MemoryModel(g.node_raw_features, g.edge_raw_features, sampler(g), 100, name, 1, 2, 0.1, 3.0, 50.0, 5.0, 70.0, 'cuda').eval()
        m.load_state_dict(deterministic_state_dict(m.state_dict(), wseed))
        m.memory_bank.__init_memory_bank__()
        return m, (lambda mm: (mm.memory_bank.node_memories.data, mm.memory_bank.node_last_updated_times.data))
    return sampler, tgat, dygformer, memory


def load_golden(name):
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name)))


# ---------------------------------------------------------------------------------------------
# Training step (train_link_prediction.py:165-257): BCE over positive and negative link probabilities, one backward pass.
def tgat_train_step(embed_fn, predictor_fn, params: dict, seed=11, start=2000, B=40, k=20):
    """embed_fn(src, dst, t, k) -> (src_emb, dst_emb) torch tensors with autograd; predictor_fn(a, b) -> (n, 1) logits;
    params: name -> tensor whose .grad is collected.  Returns {'loss': ..., 'grad.<name>': ...} as numpy."""
    g = small_graph(seed=seed)
    src, dst, t, _, neg = next(batches(g, start, 1, B))
    ps, pd = embed_fn(src, dst, t, k)
    ns, nd = embed_fn(src, neg, t, k)
    pos = predictor_fn(ps, pd).squeeze(dim=-1).sigmoid()
    negp = predictor_fn(ns, nd).squeeze(dim=-1).sigmoid()
    predicts = torch.cat([pos, negp], dim=0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
    loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
    for p in params.values():
        p.grad = None
    loss.backward()
    out = {'loss': np.asarray(loss.item(), dtype=np.float64), 'pos': pos.detach().cpu().numpy(), 'neg': negp.detach().cpu().numpy()}
    for name, p in params.items():
        out['grad.' + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return out


def predictor_template():
    from dyglib_b200.models.modules import MergeLayer
    return MergeLayer(172, 172, 172, 1)


GRAD_SAMPLES = 1024


def compact_grads(out: dict) -> dict:
    """Fixture form of a training step's outputs: every gradient tensor (keys containing 'grad.') is replaced by a strided sample
    of at most ~GRAD_SAMPLES elements ('gradc.') plus [size, sum, sum of magnitudes, largest magnitude] ('gradstat.'), so that the
    committed golden files stay small; tensors of up to GRAD_SAMPLES elements are kept whole."""
    res = {}
    for k, v in out.items():
        if 'grad.' in k and 'gradc.' not in k and 'gradstat.' not in k:
            flat = np.asarray(v, dtype=np.float32).reshape(-1)
            f64 = flat.astype(np.float64)
            stride = max(1, flat.size // GRAD_SAMPLES)
            res[k.replace('grad.', 'gradc.', 1)] = flat[::stride].copy()
            res[k.replace('grad.', 'gradstat.', 1)] = np.array([flat.size, f64.sum(), np.abs(f64).sum(), np.abs(f64).max() if flat.size else 0.0])
        else:
            res[k] = v
    return res


def assert_grads_close(got: dict, want: dict, rtol=2e-3):
    """Every gradient tensor within rtol of its own largest magnitude (fp32 accumulation order differs).  ``want`` may be in the
    compact fixture form (``compact_grads``): then the sampled elements are compared the same way and the sum / sum of
    magnitudes of the whole tensor within 10 rtol of the reference's sum of magnitudes."""
    if any(k.startswith('gradc.') for k in want):
        got = compact_grads(got)
    assert set(got) == set(want), set(got) ^ set(want)
    for key in sorted(want):
        a, b = np.asarray(got[key], dtype=np.float64), np.asarray(want[key], dtype=np.float64)
        assert a.shape == b.shape, key
        if key.startswith('gradstat.'):
            assert a[0] == b[0], key
            tol = 10 * rtol * max(b[2], 1e-7)
            assert abs(a[1] - b[1]) < tol and abs(a[2] - b[2]) < tol, (key, a, b)
            continue
        scale = max(float(np.abs(b).max()), 1e-7)
        if key.startswith('gradc.'):
            scale = max(float(want['gradstat.' + key[6:]][3]), 1e-7)
        err = float(np.abs(a - b).max()) / scale
        assert err < rtol, (key, err)


def oracle_tgat_train_step():
    """One training step of the oracle TGAT + link predictor (autograd through the oracle's torch-CPU restatement)."""
    from oracle.models import merge_layer
    _, tgat, _, _ = oracle_factories()
    g = small_graph(seed=11)
    m = tgat(g, 1)
    m.sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in m.sd.items()}
    psd = {k: v.clone().requires_grad_(True) for k, v in deterministic_state_dict(predictor_template().state_dict(), 5).items()}
    params = {'model.' + k: v for k, v in m.sd.items()}
    params.update({'pred.' + k: v for k, v in psd.items()})
    return tgat_train_step(lambda s, d, t, k: m.compute_src_dst_node_temporal_embeddings(s, d, t, k),
                           lambda a, b: merge_layer(psd, '', a, b), params)


def cuda_tgat_train_step(dropout=0.0):
    """The same step on the CUDA package (training mode, autograd through dyglib_b200/autograd.py)."""
    _, tgat, _, _ = cuda_factories()
    g = small_graph(seed=11)
    m = tgat(g, 1)
    for layer in m.temporal_conv_layers:
        layer.dropout.p = dropout
    m.train()
    pred = predictor_template().to('cuda')
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    pred.train()
    params = {'model.' + k: v for k, v in m.named_parameters()}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    return tgat_train_step(lambda s, d, t, k: m.compute_src_dst_node_temporal_embeddings(s, d, t, k), lambda a, b: pred(a, b), params)


# ---------------------------------------------------------------------------------------------
# Memory-model training step (train_link_prediction.py:236-257): `warm` batches advance the memory without gradients, then one
# batch (negative call first, then the positive call that also advances the memory) with BCE and one backward pass.
def memory_train_step(model, predictor_fn, params: dict, warm=8, B=30, k=10, seed=13):
    g = small_graph(seed=seed)
    bs = list(batches(g, 0, warm + 1, B))
    with torch.no_grad():
        for src, dst, t, eid, neg in bs[:warm]:
            model.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, k)
            model.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, k)
    src, dst, t, eid, neg = bs[warm]
    ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t, None, False, k)
    ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t, eid, True, k)
    pos = predictor_fn(ps, pd).squeeze(dim=-1).sigmoid()
    negp = predictor_fn(ns, nd).squeeze(dim=-1).sigmoid()
    predicts = torch.cat([pos, negp], dim=0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
    loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
    for p in params.values():
        p.grad = None
    loss.backward()
    out = {'loss': np.asarray(loss.item(), dtype=np.float64), 'pos': pos.detach().cpu().numpy(), 'neg': negp.detach().cpu().numpy()}
    for name, p in params.items():
        out['grad.' + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return out


def oracle_memory_train_step(name):
    from oracle.models import merge_layer
    _, _, _, memory = oracle_factories()
    m, _ = memory(small_graph(seed=13), name, 3)
    m.sd = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and 'memory_bank' not in k) else v.clone()) for k, v in m.sd.items()}
    psd = {k: v.clone().requires_grad_(True) for k, v in deterministic_state_dict(predictor_template().state_dict(), 5).items()}
    params = {'model.' + k: v for k, v in m.sd.items() if v.requires_grad and not k.startswith('embedding_module.time_encoder')}
    params.update({'pred.' + k: v for k, v in psd.items()})
    return memory_train_step(m, lambda a, b: merge_layer(psd, '', a, b), params)


def cuda_memory_train_step(name):
    _, _, _, memory = cuda_factories()
    m, _ = memory(small_graph(seed=13), name, 3)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    pred = predictor_template().to('cuda')
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    pred.train()
    params = {'model.' + k: v for k, v in m.named_parameters() if v.requires_grad}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    return memory_train_step(m, lambda a, b: pred(a, b), params)


# ---------------------------------------------------------------------------------------------
# DyGFormer training step: positive and negative pair batch, BCE, one backward pass (dropout 0 for determinism).
def dygformer_train_step(model, predictor_fn, params: dict, seed=12, start=1000, B=50):
    g = small_graph(seed=seed)
    src, dst, t, _, neg = next(batches(g, start, 1, B))
    ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t)
    ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t)
    pos = predictor_fn(ps, pd).squeeze(dim=-1).sigmoid()
    negp = predictor_fn(ns, nd).squeeze(dim=-1).sigmoid()
    predicts = torch.cat([pos, negp], dim=0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
    loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
    for p in params.values():
        p.grad = None
    loss.backward()
    out = {'loss': np.asarray(loss.item(), dtype=np.float64), 'pos': pos.detach().cpu().numpy(), 'neg': negp.detach().cpu().numpy()}
    for name, p in params.items():
        out['grad.' + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return out


DYG_TRAIN_CASES = ((4, 32),)


def oracle_dygformer_train_step(P, L):
    from oracle.models import merge_layer
    _, _, dygformer, _ = oracle_factories()
    m = dygformer(small_graph(seed=12), P, L, 2)
    m.sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in m.sd.items()}
    psd = {k: v.clone().requires_grad_(True) for k, v in deterministic_state_dict(predictor_template().state_dict(), 5).items()}
    params = {'model.' + k: v for k, v in m.sd.items()}
    params.update({'pred.' + k: v for k, v in psd.items()})
    return dygformer_train_step(m, lambda a, b: merge_layer(psd, '', a, b), params)


def cuda_dygformer_train_step(P, L):
    _, _, dygformer, _ = cuda_factories()
    m = dygformer(small_graph(seed=12), P, L, 2)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    m.train()
    pred = predictor_template().to('cuda')
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    pred.train()
    params = {'model.' + k: v for k, v in m.named_parameters()}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    return dygformer_train_step(m, lambda a, b: pred(a, b), params)


# ---------------------------------------------------------------------------------------------
# GraphMixer (a caller of the path): eval embeddings of two batches and one training step.
def run_graphmixer_cases(model, k=20, time_gap=300):
    out = {}
    g = small_graph(seed=14)
    with torch.no_grad():
        for bi, (src, dst, t, _, neg) in enumerate(batches(g, 1500, 2, 40)):
            for tag, d in (('pos', dst), ('neg', neg)):
                a, b = model.compute_src_dst_node_temporal_embeddings(src, d, t, k, time_gap)
                out[f'eval_{bi}_{tag}_src'], out[f'eval_{bi}_{tag}_dst'] = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    return out


def graphmixer_train_step(model, predictor_fn, params: dict, k=20, time_gap=300, B=40):
    g = small_graph(seed=14)
    src, dst, t, _, neg = next(batches(g, 2000, 1, B))
    ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t, k, time_gap)
    ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t, k, time_gap)
    pos = predictor_fn(ps, pd).squeeze(dim=-1).sigmoid()
    negp = predictor_fn(ns, nd).squeeze(dim=-1).sigmoid()
    predicts = torch.cat([pos, negp], dim=0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
    loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
    for p in params.values():
        p.grad = None
    loss.backward()
    out = {'loss': np.asarray(loss.item(), dtype=np.float64), 'pos': pos.detach().cpu().numpy(), 'neg': negp.detach().cpu().numpy()}
    for name, p in params.items():
        out['grad.' + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return out


def graphmixer_template():
    from dyglib_b200.models.GraphMixer import GraphMixer
    return GraphMixer


def oracle_graphmixer(train=False):
    from oracle.sampler import OracleSampler
    from oracle.models import OracleGraphMixer
    g = small_graph(seed=14)
    sd = deterministic_state_dict(graphmixer_template()(g.node_raw_features, g.edge_raw_features, None, 100, 20, 2, device='cpu').state_dict(), 6)
    if train:
        sd = {k: v.clone().requires_grad_(not k.startswith('time_encoder')) for k, v in sd.items()}
    samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
    return OracleGraphMixer(sd, g.node_raw_features, g.edge_raw_features, samp, 2)


def oracle_graphmixer_train_step():
    from oracle.models import merge_layer
    m = oracle_graphmixer(train=True)
    psd = {k: v.clone().requires_grad_(True) for k, v in deterministic_state_dict(predictor_template().state_dict(), 5).items()}
    params = {'model.' + k: v for k, v in m.sd.items() if v.requires_grad}
    params.update({'pred.' + k: v for k, v in psd.items()})
    return graphmixer_train_step(m, lambda a, b: merge_layer(psd, '', a, b), params)


def cuda_graphmixer(train=False):
    from dyglib_b200.utils.utils import get_neighbor_sampler
    g = small_graph(seed=14)
    m = graphmixer_template()(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(g, 'recent'), 100, 20, 2,
                              dropout=0.0 if train else 0.1, device='cuda')
    m.load_state_dict(deterministic_state_dict(m.state_dict(), 6))
    return m.train() if train else m.eval()


def cuda_graphmixer_train_step():
    m = cuda_graphmixer(train=True)
    pred = predictor_template().to('cuda')
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    pred.train()
    params = {'model.' + k: v for k, v in m.named_parameters() if v.requires_grad}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    return graphmixer_train_step(m, lambda a, b: pred(a, b), params)


# ---------------------------------------------------------------------------------------------
# TCL (a caller of the path): eval embeddings of two batches and one training step (k = 20: 21 depths).
def run_tcl_cases(model, k=20):
    out = {}
    g = small_graph(seed=15)
    with torch.no_grad():
        for bi, (src, dst, t, _, neg) in enumerate(batches(g, 1500, 2, 40)):
            for tag, d in (('pos', dst), ('neg', neg)):
                a, b = model.compute_src_dst_node_temporal_embeddings(src, d, t, k)
                out[f'eval_{bi}_{tag}_src'], out[f'eval_{bi}_{tag}_dst'] = a.detach().cpu().numpy(), b.detach().cpu().numpy()
    return out


def tcl_train_step(model, predictor_fn, params: dict, k=20, B=40):
    g = small_graph(seed=15)
    src, dst, t, _, neg = next(batches(g, 2000, 1, B))
    ps, pd = model.compute_src_dst_node_temporal_embeddings(src, dst, t, k)
    ns, nd = model.compute_src_dst_node_temporal_embeddings(src, neg, t, k)
    pos = predictor_fn(ps, pd).squeeze(dim=-1).sigmoid()
    negp = predictor_fn(ns, nd).squeeze(dim=-1).sigmoid()
    predicts = torch.cat([pos, negp], dim=0)
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(negp)], dim=0)
    loss = torch.nn.functional.binary_cross_entropy(predicts, labels)
    for p in params.values():
        p.grad = None
    loss.backward()
    out = {'loss': np.asarray(loss.item(), dtype=np.float64), 'pos': pos.detach().cpu().numpy(), 'neg': negp.detach().cpu().numpy()}
    for name, p in params.items():
        out['grad.' + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return out


def tcl_template():
    from dyglib_b200.models.TCL import TCL
    return TCL


def oracle_tcl(train=False):
    from oracle.sampler import OracleSampler
    from oracle.models import OracleTCL
    g = small_graph(seed=15)
    sd = deterministic_state_dict(tcl_template()(g.node_raw_features, g.edge_raw_features, None, 100, 2, 2, 21, device='cpu').state_dict(), 7)
    if train:
        sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    samp = OracleSampler(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, g.num_nodes, 'recent')
    return OracleTCL(sd, g.node_raw_features, g.edge_raw_features, samp, 2, 2)


def oracle_tcl_train_step():
    from oracle.models import merge_layer
    m = oracle_tcl(train=True)
    psd = {k: v.clone().requires_grad_(True) for k, v in deterministic_state_dict(predictor_template().state_dict(), 5).items()}
    params = {'model.' + k: v for k, v in m.sd.items()}
    params.update({'pred.' + k: v for k, v in psd.items()})
    return tcl_train_step(m, lambda a, b: merge_layer(psd, '', a, b), params)


def cuda_tcl(train=False):
    from dyglib_b200.utils.utils import get_neighbor_sampler
    g = small_graph(seed=15)
    m = tcl_template()(g.node_raw_features, g.edge_raw_features, get_neighbor_sampler(g, 'recent'), 100, 2, 2, 21,
                       dropout=0.0 if train else 0.1, device='cuda')
    m.load_state_dict(deterministic_state_dict(m.state_dict(), 7))
    return m.train() if train else m.eval()


def cuda_tcl_train_step():
    m = cuda_tcl(train=True)
    pred = predictor_template().to('cuda')
    pred.load_state_dict(deterministic_state_dict(pred.state_dict(), 5))
    pred.train()
    params = {'model.' + k: v for k, v in m.named_parameters()}
    params.update({'pred.' + k: v for k, v in pred.named_parameters()})
    return tcl_train_step(m, lambda a, b: pred(a, b), params)
