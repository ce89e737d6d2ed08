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
    """One context for the whole GPU session.  Fails loudly if the CUDA library is absent.
    Run the suite with ASLAM_B200_LIB=<repo>/aria_slam_b200/libaslam_b200_dbg.so and every staged index of every kernel is
    checked on the device during ALL GPU tests; the session then fails here if any check tripped."""
    import ctypes as C
    from aria_slam_b200._lib import Context, lib
    c = Context(0)
    yield c
    v = C.c_uint64(0)
    debug_build = lib().aslam_debug_bounds(C.byref(v))
    c.close()
    if debug_build:
        print("\n[aslam] bounds-checking build: %d violations over the session" % v.value)
        assert v.value == 0, "the debug build counted %d out-of-bounds staged indices" % v.value


def kp_keys(k):
    """Keypoint identity used by every set comparison: (octave, x, y) as produced (exact floats)."""
    return list(zip(k["octave"].tolist(), k["x"].tolist(), k["y"].tolist()))
