import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def decoder():
    """One Decoder (gst_ctx) per test session; fails loudly when there is no sm_100 device."""
    import gst_b200
    dec = gst_b200.Decoder(0)
    yield dec
    dec.close()


@pytest.fixture(scope="session")
def ref_lib():
    import gst_fixtures as fx
    L = fx.ref()
    if L is None:
        pytest.skip("oracle/_ref/libgst_ref.so not built (needs /root/reference)")
    return L
