import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx():
    """A libflame_b200 context on cuda:0 with 8 slots. Fails loudly without a GPU / the .so."""
    from flame_b200 import _lib
    ctx = _lib.Context(0, 8)
    yield ctx
    ctx.close()
