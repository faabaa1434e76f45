import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The reference's own hot path (oracle/_ref/libsmref.so)."""
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libsmref.so not built (make -C oracle ref)")
    return refapi.get()
