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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One context for the whole GPU session.  Fails loudly if the CUDA library is absent."""
    from aria_slam_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def kp_keys(k):
    """Keypoint identity used by every set comparison: (octave, x, y) as produced (exact floats)."""
    return list(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
