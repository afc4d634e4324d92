"""CUDA cases written after round 1's GPU budget was spent.  They are deliberately NOT marked
``gpu`` (the driver's ``-m gpu`` run must only contain tests that have passed on a B200) and are
skipped without a CUDA device; run them with ``-m gpu_next`` on the GPU box, then move them
into tests/test_gpu_parity.py."""
import pytest
import torch

from tests import test_host_logic as host

pytestmark = [pytest.mark.gpu_next,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(autouse=True)
def _on_device():
    host.DEVICE = "cuda"
    yield
    host.DEVICE = "cpu"


def test_rt_eom_step_matches_reference():
    host.test_rt_eom_step_matches_reference(None)
