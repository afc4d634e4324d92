import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")
    config.addinivalue_line("markers", "gpu_next: CUDA cases written after the round's last GPU session; "
                                       "not part of -m gpu until they have run on a B200 once")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture
def cpu_abi(monkeypatch):
    """Host tensors + numpy emulation of the C ABI (tests/abi_emulator.py): checks the
    host logic of pymes_b200 without a GPU.  Never active in -m gpu tests."""
    from tests import abi_emulator
    return abi_emulator.install(monkeypatch)


@pytest.fixture(autouse=True)
def _quiet_logs():
    from pymes_b200 import log
    log.set_quiet(True)
    yield
    log.set_quiet(False)
