import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def covt():
    """The product package (ctypes binding over libcovt_b200.so)."""
    import covt_loader
    return covt_loader.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def gen():
    from tools.gen import gen as G
    G.lib()
    return G


@pytest.fixture(scope="session")
def decoder(covt):
    """A live decoder context on cuda:0 — fails loudly (no CPU fallback) if the library or GPU is missing."""
    covt.build()
    return covt.Decoder(0)


@pytest.fixture(scope="session")
def fixtures():
    import util
    return util.load_fixture_tiles()
