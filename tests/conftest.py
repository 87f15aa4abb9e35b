import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.vrdd_oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return dict(np.load(path))


@pytest.fixture()
def renderer():
    """A fresh vrdd handle on cuda:0.  No fallback: fails if the library or the GPU is missing."""
    import vrdd_b200 as V
    r = V.Renderer(0)
    yield r
    r.close()
