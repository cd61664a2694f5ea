import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cm():
    """The product package (ctypes binding of libcm31.so)."""
    return importlib.import_module("cairo-m_b200")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from tests import oracle_lib
    return oracle_lib
