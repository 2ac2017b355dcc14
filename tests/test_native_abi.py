"""CPU: libdygb200.so builds for sm_100a, loads, and exports every symbol include/dygb200.h declares."""
import os
import re

from dyglib_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'dygb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dyg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/dygb200.h but not exported'
    assert lib.dyg_abi_version() == _native.ABI_VERSION


def test_python_signatures_cover_header():
    declared = set(declared_symbols()) - {'dyg_last_error', 'dyg_abi_version', 'dyg_ln_ffn_workspace_bytes', 'dyg_csr_fence_entries', 'dyg_tgn_step_sizeof',
                                         'dyg_attn_block_workspace_bytes', 'dyg_radix_sort_workspace_entries'}   # non-int returns
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)


def test_header_argument_counts_match_ctypes():
    text = open(os.path.join(ROOT, 'include', 'dygb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    for name, args in _native.SIGNATURES.items():
        m = re.search(r'\bint\s+' + name + r'\s*\((.*?)\)\s*;', text, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(',') if a.strip()]) == len(args), name


def test_no_cpu_fallback():
    import pytest
    import torch
    from dyglib_b200 import ops
    if torch.cuda.is_available():
        pytest.skip('has a GPU')
    with pytest.raises(RuntimeError):
        ops.time_encode(torch.zeros(4), torch.ones(8), torch.zeros(8))
    from dyglib_b200.utils.utils import NeighborSampler
    with pytest.raises(RuntimeError):
        NeighborSampler([[], [(2, 1, 1.0)], [(1, 1, 1.0)]], 'recent')
