import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libfrx.so and the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as ge

    ge.build()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import binding

    return binding
