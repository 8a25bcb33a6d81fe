import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine. No fallback: on a box without a usable GPU this errors, it does not skip."""
    from rts_b200 import lib
    eng = lib.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library and the oracle are built (both build without a GPU)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "rts_b200", "librts_b200.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "rts_b200", "csrc")], check=True, capture_output=True)
    import oracle_api
    oracle_api.oracle()
