import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import lib
    lib.load()
    return lib


@pytest.fixture(scope="session")
def engine():
    """One GPU context for the whole session; fails loudly when the CUDA library or a GPU is missing."""
    from snacc_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
