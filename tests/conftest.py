import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), compiled on demand from oracle/mixlab_oracle.c."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def mxl():
    """The product library through its ctypes view; built in-tree if missing (nvcc needs no GPU)."""
    import mixlab_b200
    from mixlab_b200 import build as mxl_build
    if not os.path.exists(mixlab_b200.api.LIB_PATH):
        mxl_build.build()
    mixlab_b200.lib()
    return mixlab_b200


@pytest.fixture()
def ctx48(mxl):
    c = mxl.Context(device=0, sample_rate=48000, samples_per_tick=800)
    yield c
    c.close()


@pytest.fixture()
def ctx44(mxl):
    c = mxl.Context(device=0, sample_rate=44100, samples_per_tick=735)
    yield c
    c.close()
