"""Rewrite the training-step fixtures under tests/golden/ in the compact form of tests/helpers.compact_grads (strided samples + sums
of every gradient tensor instead of the full tensors).  The make_golden_*.py scripts write this form directly; this script converts
files written before that:
    python scripts/compact_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import compact_grads  # noqa: E402

for name in ('tgat_train.npz', 'memory_train.npz', 'dygformer_train.npz', 'graphmixer.npz', 'tcl.npz'):
    path = os.path.join(ROOT, 'tests', 'golden', name)
    d = dict(np.load(path))
    before = os.path.getsize(path)
    np.savez_compressed(path, **compact_grads(d))
    print(name, before, '->', os.path.getsize(path))
