import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def bmp():
    return open(os.path.join(ROOT, "tests", "golden", "Test.bmp"), "rb").read()


@pytest.fixture(scope="session")
def test_lz():
    return open(os.path.join(ROOT, "tests", "golden", "Test.lz"), "rb").read()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def codec():
    """The product: libaurora_cuda.so on the visible B200(s).  No fallback: absence is an error."""
    import __graft_entry__ as g
    from auroralib.compression_b200 import BatchCodec, _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    c = BatchCodec()
    yield c
    c.close()
