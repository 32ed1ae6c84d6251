import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle binding (test infrastructure; built on demand with oracle/Makefile)."""
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def xnart_lib():
    """The product library; GPU tests must fail loudly if it is missing (no CPU fallback)."""
    from xna_ray_trace_b200 import capi

    return capi
