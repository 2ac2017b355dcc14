"""CPU: the host-side pieces either side of the hot path against the unmodified reference where it
is importable here (/root/reference, build container) and against scikit-learn."""
import os
import random
import sys

import numpy as np
import pytest
import torch

REF = '/root/reference'


def _write_dataset(root, name, E=4000, nu=120, ni=60, feat=8, seed=0):
    import pandas as pd
    rng = np.random.default_rng(seed)
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    u = rng.integers(1, nu + 1, E)
    i = rng.integers(nu + 1, nu + ni + 1, E)
    ts = np.sort(rng.integers(0, 100000, E)).astype(np.float64)
    pd.DataFrame({'u': u, 'i': i, 'ts': ts, 'label': rng.integers(0, 2, E), 'idx': np.arange(1, E + 1)}).to_csv(
        os.path.join(d, f'ml_{name}.csv'))
    np.save(os.path.join(d, f'ml_{name}.npy'), rng.standard_normal((E + 1, feat)))
    np.save(os.path.join(d, f'ml_{name}_node.npy'), np.zeros((nu + ni + 1, feat)))


def _same_data(a, b):
    for f in ('src_node_ids', 'dst_node_ids', 'node_interact_times', 'edge_ids', 'labels'):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and np.array_equal(x, y), f
    assert a.num_interactions == b.num_interactions and a.unique_node_ids == b.unique_node_ids


def test_link_prediction_split_properties(tmp_path):
    from dyglib_b200.utils.DataLoader import get_link_prediction_data, get_idx_data_loader
    _write_dataset(str(tmp_path), 'toy')
    nf, ef, full, train, val, test, nn_val, nn_test = get_link_prediction_data('toy', 0.15, 0.15, root=str(tmp_path), verbose=False)
    assert nf.shape[1] == 172 and ef.shape[1] == 172 and np.all(ef[:, 8:] == 0)
    assert val.num_interactions + test.num_interactions + int((full.node_interact_times <= val.node_interact_times.min() - 1e-9).sum()) \
        <= full.num_interactions
    assert train.node_interact_times.max() < val.node_interact_times.min() <= val.node_interact_times.max() < test.node_interact_times.min()
    new_nodes = full.unique_node_ids - train.unique_node_ids
    for d in (nn_val, nn_test):
        assert all((s in new_nodes) or (t in new_nodes) for s, t in zip(d.src_node_ids, d.dst_node_ids))
    batches = [b.numpy() for b in get_idx_data_loader(list(range(10)), batch_size=4, shuffle=False)]
    assert [len(b) for b in batches] == [4, 4, 2] and batches[0][0] == 0


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference tree only exists in the build container')
def test_link_prediction_data_matches_reference(tmp_path, monkeypatch):
    """Same files, same ratios: every returned array equals the reference loader's.  The reference calls random.sample(set),
    which Python >= 3.11 rejects; it is given the Python <= 3.10 behaviour (population = tuple(set)) for this comparison."""
    from dyglib_b200.utils.DataLoader import get_link_prediction_data, get_node_classification_data
    _write_dataset(str(tmp_path / 'processed_data'), 'toy', seed=3)
    monkeypatch.chdir(tmp_path)
    sys.path.insert(0, REF)
    try:
        import importlib
        ref = importlib.import_module('utils.DataLoader')
    finally:
        sys.path.remove(REF)
    real_sample = random.sample
    monkeypatch.setattr(random, 'sample', lambda pop, k: real_sample(tuple(pop) if isinstance(pop, (set, frozenset)) else pop, k))
    want = ref.get_link_prediction_data('toy', 0.15, 0.15)
    got = get_link_prediction_data('toy', 0.15, 0.15, verbose=False)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    for a, b in zip(got[2:], want[2:]):
        _same_data(a, b)
    want = ref.get_node_classification_data('toy', 0.15, 0.15)
    got = get_node_classification_data('toy', 0.15, 0.15)
    for a, b in zip(got[2:], want[2:]):
        _same_data(a, b)


def _neg_sampler_cls():
    # the class lives next to the device sampler; importing the module does not need a GPU
    from dyglib_b200.utils.utils import NegativeEdgeSampler
    return NegativeEdgeSampler


def test_negative_edge_sampler_random_stream():
    cls = _neg_sampler_cls()
    rng = np.random.default_rng(1)
    src, dst = rng.integers(1, 50, 500), rng.integers(50, 90, 500)
    s = cls(src, dst, seed=7)
    a = [s.sample(40) for _ in range(3)]
    rs = np.random.RandomState(7)
    us, ud = np.unique(src), np.unique(dst)
    for ns, nd in a:
        assert np.array_equal(ns, us[rs.randint(0, len(us), 40)])
        assert np.array_equal(nd, ud[rs.randint(0, len(ud), 40)])
    s.reset_random_state()
    assert np.array_equal(s.sample(40)[1], a[0][1])
    with pytest.raises(ValueError):
        cls(src, dst, negative_sample_strategy='nope', seed=1).sample(3)


@pytest.mark.skipif(not os.path.isdir(REF), reason='the reference tree only exists in the build container')
def test_negative_edge_sampler_matches_reference():
    sys.path.insert(0, REF)
    try:
        import importlib
        ref = importlib.import_module('utils.utils')
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(2)
    src, dst = rng.integers(1, 80, 900), rng.integers(80, 140, 900)
    a, b = _neg_sampler_cls()(src, dst, seed=3), ref.NegativeEdgeSampler(src, dst, seed=3)
    for size in (1, 200, 37):
        x, y = a.sample(size), b.sample(size)
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[0].dtype == y[0].dtype


@pytest.mark.parametrize('ties', [False, True])
def test_link_metrics_match_sklearn(ties):
    from sklearn.metrics import average_precision_score, roc_auc_score
    from dyglib_b200.utils.metrics import get_link_prediction_metrics, link_prediction_metrics_tensors
    g = torch.Generator().manual_seed(5)
    for n in (2, 17, 400, 5000):
        labels = torch.cat([torch.ones(n), torch.zeros(n)])
        p = torch.rand(2 * n, generator=g)
        if ties:
            p = torch.round(p * 7) / 7
        got = get_link_prediction_metrics(p, labels)
        assert abs(got['average_precision'] - average_precision_score(labels.numpy(), p.numpy())) < 1e-12
        assert abs(got['roc_auc'] - roc_auc_score(labels.numpy(), p.numpy())) < 1e-12
        ap, auc = link_prediction_metrics_tensors(p, labels)
        assert ap.dtype == torch.float64 and ap.dim() == 0 and auc.dim() == 0


def test_every_source_file_compiles():
    """A syntax error in a module that only the GPU tests import would otherwise surface on the GPU box."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [f for pat in ('dyglib_b200/**/*.py', 'oracle/*.py', 'tests/*.py', 'scripts/*.py', '*.py') for f in glob.glob(os.path.join(root, pat), recursive=True)]
    assert len(files) > 20
    for f in files:
        compile(open(f).read(), f, 'exec')
