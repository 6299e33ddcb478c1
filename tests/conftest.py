import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def cuda_lib():
    """The built C-ABI library on a CUDA device; GPU tests fail (not skip) if it is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    from soft_contrastive_learning_b200 import _lib
    L = _lib.lib()
    assert L.scl_device_ok() == 0, "device is not sm_100"
    return L
