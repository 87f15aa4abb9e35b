import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


_ORACLE = []


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle with the rounding of the reference's own build (FMA pattern of its PTX + the B200's
    rsqrt.approx table, Oracle.set_reference_build) — the default of libvrdd.so, so ray geometry is bit-identical
    on both sides.  The rounding is re-applied before every test (_oracle_rounding); a test that wants the source's
    uncontracted order says oracle.set_reference_build(False)."""
    from oracle.vrdd_oracle import Oracle
    o = Oracle()
    o.set_reference_build(True)
    _ORACLE.append(o)
    return o


@pytest.fixture(autouse=True)
def _oracle_rounding():
    for o in _ORACLE:
        o.set_reference_build(True)
    yield


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v2.npz")
    return dict(np.load(path))


@pytest.fixture()
def renderer():
    """A fresh vrdd handle on cuda:0.  No fallback: fails if the library or the GPU is missing."""
    import vrdd_b200 as V
    r = V.Renderer(0)
    yield r
    r.close()
