import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    return ge.load_package()


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r

    r.lib()
    return r


@pytest.fixture(scope="session")
def hostsim():
    """The kernels' stage bodies compiled for the host (tests/hostsim) -- logic check without a GPU."""
    import ctypes

    import __graft_entry__ as ge

    so = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    csrc = os.path.join(ROOT, "esp32-opus-player_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc)])
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        ge.build_hostsim()
    H = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    H.hs_celt_parse.argtypes = [P, I, I, I, I, I, P, P, P]
    H.hs_celt_new.restype = P
    H.hs_celt_new.argtypes = [I]
    H.hs_celt_free.argtypes = [P]
    H.hs_celt_decode.argtypes = [P, P, I, I, I, I, P, P]
    H.hs_opus_new.restype = P
    H.hs_opus_new.argtypes = [I]
    H.hs_opus_free.argtypes = [P]
    H.hs_opus_decode.argtypes = [P, I, P, I, P]
    return H
