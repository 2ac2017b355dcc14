"""CPU: the oracle reproduces the golden vectors the unmodified reference produced (scripts/make_golden.py)."""
import numpy as np
import pytest

from helpers import run_sampler_cases, run_model_cases, oracle_factories, load_golden
from oracle.sampler import pad_sequences, count_nodes_appearances


def test_oracle_sampler_matches_golden():
    sampler, _, _, _ = oracle_factories()
    got = run_sampler_cases(lambda g, st, seed, tsf: sampler(g, st, seed, tsf),
                            lambda s, g, nodes, times, lists, P, L: pad_sequences(nodes, times, lists[0], lists[1], lists[2], P, L),
                            count_nodes_appearances)
    gold = load_golden('sampler.npz')
    assert set(got) == set(gold)
    for k in gold:
        assert got[k].dtype == gold[k].dtype, k
        assert np.array_equal(got[k], gold[k]), k


@pytest.mark.parametrize('which', ['tgat', 'dygformer', 'TGN', 'DyRep', 'JODIE'])
def test_oracle_models_match_golden(which):
    _, tgat, dygformer, memory = oracle_factories()
    got = run_model_cases(tgat, dygformer, memory, which=(which,))
    gold = load_golden('models.npz')
    assert len(got) > 0
    for k in got:
        np.testing.assert_allclose(got[k], gold[k], rtol=1e-4, atol=1e-5, err_msg=k)


def test_oracle_tgat_training_step_matches_golden():
    """Loss and every parameter gradient of one training step (train_link_prediction.py:165-257) against the reference's
    (scripts/make_golden_train.py)."""
    from helpers import oracle_tgat_train_step, assert_grads_close
    got = oracle_tgat_train_step()
    gold = load_golden('tgat_train.npz')
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    assert_grads_close(got, gold, rtol=1e-3)


@pytest.mark.parametrize('name', ['TGN', 'DyRep', 'JODIE'])
def test_oracle_memory_training_step_matches_golden(name):
    """Memory models: 8 batches advance the memory without gradients, then one training batch (negative call, positive call,
    BCE, backward); loss and every parameter gradient against the reference's (scripts/make_golden_train.py)."""
    from helpers import oracle_memory_train_step, assert_grads_close
    got = oracle_memory_train_step(name)
    gold = {k[len(name) + 1:]: v for k, v in load_golden('memory_train.npz').items() if k.startswith(name + '.')}
    np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
    assert_grads_close(got, gold, rtol=1e-3)


def test_oracle_dygformer_training_step_matches_golden():
    """DyGFormer (patch 4, length 32): loss and every parameter gradient of one training step against the reference's."""
    from helpers import oracle_dygformer_train_step, assert_grads_close, DYG_TRAIN_CASES
    for P, L in DYG_TRAIN_CASES:
        got = oracle_dygformer_train_step(P, L)
        gold = {k[len(f'P{P}_L{L}.'):]: v for k, v in load_golden('dygformer_train.npz').items() if k.startswith(f'P{P}_L{L}.')}
        np.testing.assert_allclose(got['loss'], gold['loss'], rtol=1e-5)
        assert_grads_close(got, gold, rtol=1e-3)


def test_oracle_graphmixer_matches_golden():
    """GraphMixer (a caller of the path): eval embeddings of two batches and one training step against the reference's
    (scripts/make_golden_graphmixer.py)."""
    from helpers import oracle_graphmixer, oracle_graphmixer_train_step, run_graphmixer_cases, assert_grads_close
    gold = load_golden('graphmixer.npz')
    got = run_graphmixer_cases(oracle_graphmixer())
    for k in got:
        np.testing.assert_allclose(got[k], gold[k], rtol=1e-4, atol=1e-5, err_msg=k)
    tr = oracle_graphmixer_train_step()
    gtr = {k[len('train.'):]: v for k, v in gold.items() if k.startswith('train.')}
    np.testing.assert_allclose(tr['loss'], gtr['loss'], rtol=1e-5)
    assert_grads_close(tr, gtr, rtol=1e-3)


def test_oracle_tcl_matches_golden():
    """TCL (a caller of the path): eval embeddings of two batches and one training step against the reference's
    (scripts/make_golden_tcl.py)."""
    from helpers import oracle_tcl, oracle_tcl_train_step, run_tcl_cases, assert_grads_close
    gold = load_golden('tcl.npz')
    got = run_tcl_cases(oracle_tcl())
    for k in got:
        np.testing.assert_allclose(got[k], gold[k], rtol=1e-4, atol=1e-5, err_msg=k)
    tr = oracle_tcl_train_step()
    gtr = {k[len('train.'):]: v for k, v in gold.items() if k.startswith('train.')}
    np.testing.assert_allclose(tr['loss'], gtr['loss'], rtol=1e-5)
    assert_grads_close(tr, gtr, rtol=1e-3)
