import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_libs():
    """Both product libraries, on a machine that has a GPU.  Fails (not skips) if the extension is missing."""
    import numpy as np
    from cmfrec_b200 import _lib
    libs = {np.dtype(np.float32): _lib.load(np.float32), np.dtype(np.float64): _lib.load(np.float64)}
    assert libs[np.dtype(np.float32)].cmfb200_device_count() >= 1, "no CUDA device visible"
    return libs
