import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as _orc

    _orc.build()
    return _orc


@pytest.fixture(autouse=True)
def _graph_cache_off_by_default(request):
    """GPU tests compare library options on repeated calls with the same tensors; rw.walk's graph
    cache would keep the first preparation.  Tests of the cache switch it on themselves."""
    if "gpu" not in request.keywords:
        yield
        return
    from torch_random_walk_b200 import native

    native.set_graph_cache(False)
    yield
    native.set_graph_cache(False)
    native.reset_options()  # a test that changed library options must not pass them on to the next one
