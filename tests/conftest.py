import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand here; on the GPU box the prebuilt .so travels)."""
    from infur_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib.load()


@pytest.fixture(scope="session")
def handle(lib):
    from infur_b200 import processors as P

    h = P.Handle(device=0, max_batch=8, compute_aux=True, blend=True)
    yield h
    h.close()


@pytest.fixture(scope="session")
def tiny():
    from infur_b200 import synth

    return synth.ensure_fixture("fcn_tiny")


@pytest.fixture(scope="session")
def fcn50():
    from infur_b200 import synth

    return synth.ensure_fixture("fcn50")
