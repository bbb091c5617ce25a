import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "hostsim")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import cmpc_loader
    return cmpc_loader.load()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    class Lazy(dict):
        def __missing__(self, N):
            name = "golden_N%d.npz" % N if isinstance(N, int) else "golden_%s.npz" % N
            self[N] = dict(np.load(os.path.join(ROOT, "tests", "golden", name)))
            return self[N]
    return Lazy()


@pytest.fixture(scope="session")
def walk_ticks():
    import numpy as np
    class Lazy(dict):
        def __missing__(self, N):
            self[N] = dict(np.load(os.path.join(ROOT, "tests", "golden", "walk_ticks_N%d.npz" % N)))
            return self[N]
    return Lazy()
