import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctx_factory():
    """Creates device contexts through the C ABI; GPU tests only."""
    from mom6_b200.api import Context
    made = []

    def make(dom, device=0):
        c = Context(dom, device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
