import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("bulklmm.jl_b200", "oracle"):
    path = os.path.join(ROOT, sub)
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One blmm context on cuda:0.  No fallback: a missing library or GPU is an error."""
    from blmm_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
