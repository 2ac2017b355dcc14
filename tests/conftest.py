import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the upstream reference importable at /root/reference')


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_DIR, 'models'))


def import_reference():
    """Import the unmodified reference (only present in the build container)."""
    if not have_reference():
        pytest.skip('reference tree not present (GPU box)')
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import importlib
    mods = {}
    mods['utils'] = importlib.import_module('utils.utils')
    mods['DataLoader'] = importlib.import_module('utils.DataLoader')
    mods['modules'] = importlib.import_module('models.modules')
    mods['TGAT'] = importlib.import_module('models.TGAT')
    mods['DyGFormer'] = importlib.import_module('models.DyGFormer')
    mods['MemoryModel'] = importlib.import_module('models.MemoryModel')
    mods['GraphMixer'] = importlib.import_module('models.GraphMixer')
    mods['TCL'] = importlib.import_module('models.TCL')
    return mods


@pytest.fixture(scope='session')
def ref():
    return import_reference()
