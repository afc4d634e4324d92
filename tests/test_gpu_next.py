"""CUDA cases that were written after round 1's last GPU session.  They are deliberately NOT
marked ``gpu`` (the driver's ``-m gpu`` run must only contain tests that have passed on a B200)
and are skipped without a CUDA device; run them with ``-m gpu_next`` on the GPU box, then move
them into tests/test_gpu_parity.py."""
import pytest
import torch

from tests import test_host_logic as host

pytestmark = [pytest.mark.gpu_next,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


@pytest.fixture(autouse=True)
def _on_device():
    host.DEVICE, host.RUN_PENDING_GPU_CASES = "cuda", True
    yield
    host.DEVICE, host.RUN_PENDING_GPU_CASES = "cpu", False


def test_generated_operand_role_swap_and_single_index_contractions():
    host.test_ueg_virtual_block_descriptor(None)


def test_eom_with_never_materialised_abcd():
    host.test_eom_with_never_materialised_abcd(None)
