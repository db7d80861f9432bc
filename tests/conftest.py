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
    # The C-ABI tests need libblmm_b200.so: build it (nvcc cross-compiles without a GPU) if this checkout has not
    # been built yet.  Stale objects are rebuilt, so the tests always run against the sources they sit next to.
    import importlib.util
    spec = importlib.util.spec_from_file_location("blmm_build", os.path.join(ROOT, "bulklmm.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        mod.build_library()


@pytest.fixture(scope="session")
def engine():
    """One blmm context on cuda:0.  No fallback: a missing library or GPU is an error."""
    from blmm_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
