import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/librcv_oracle.so, built on demand with gcc)."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_golden.npz"))


@pytest.fixture(scope="session")
def rcv():
    """The product: rustcv_b200 over librcv_imgproc.so, bound to cuda:0.  No fallback:
    if the library is missing or the GPU is not a B200 this fixture errors out."""
    import rustcv_b200

    rustcv_b200.imgproc.init(0)
    return rustcv_b200
