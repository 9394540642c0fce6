import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _inference_tests_run_without_autograd(request):
    """With gradients enabled the set models take the training path (fp32 forward that keeps activations); the parity
    tests of the inference kernels therefore run under no_grad, as the reference's eval loops do (Code/pceval.py:88).
    tests/test_gpu_train.py and tests that ask for a backward pass opt out."""
    import torch
    name = request.module.__name__
    if name.endswith("test_gpu_train") or "backward" in request.node.name:
        yield
        return
    with torch.no_grad():
        yield
