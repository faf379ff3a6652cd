"""pytest configuration: registers the `gpu` marker and makes the repo root importable.

`-m "not gpu"` covers the oracle against the committed golden vectors, the host logic and the
C-ABI symbol table; `-m gpu` tests are the parity tests proper and need a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:                                     # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def lib_built():
    """Build libvlnimagine.so if it is stale (nvcc cross-compiles without a GPU)."""
    import importlib
    build = importlib.import_module('vln_imagine_b200.build')
    return build.build()
