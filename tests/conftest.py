import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gcs():
    return importlib.import_module("2d_geometry_constraint_solver_b200")


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build_cuda()
    g.build_oracle()
    return g


@pytest.fixture(scope="session")
def gpu(gcs, built):
    """Initialised CUDA library; the GPU tests FAIL (not skip) if the extension cannot run."""
    gcs.capi.init([0])
    return gcs.capi
