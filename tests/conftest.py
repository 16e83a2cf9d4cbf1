import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    # a clean checkout has no built artefacts (they are git-ignored): build the C-ABI library once (nvcc cross-compiles on CPU)
    if not os.path.exists(os.path.join(g.PKG_DIR, "csrc", "librbo.so")) and not os.environ.get("RBO_LIB_PATH"):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def oracle_problem(orc, wl, sur, rn, starts, mode=1, **kw):
    import __graft_entry__ as g
    return g._oracle_problem(orc, wl, sur, rn, starts, mode, **kw)


def relerr(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)))) if a.size else 0.0
