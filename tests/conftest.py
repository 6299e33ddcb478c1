import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)
GOLDEN = os.path.join(ROOT, "tests", "golden")
_MEASURED = {}


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


@pytest.fixture
def tune(cuda_lib):
    """Kernel-selection knobs of the library (scl_set_tuning): ``tune("SCL_WMS_STREAM", 1)``; restored after the test.
    The library reads its SCL_* environment variables once at load time, so tests go through the C ABI instead."""
    from soft_contrastive_learning_b200 import _lib
    saved = {}

    def set_(name, value):
        if name not in saved:
            saved[name] = _lib.get_tuning(name)
        _lib.set_tuning(name, value)

    yield set_
    for k, v in saved.items():
        _lib.set_tuning(k, v)


@pytest.fixture
def measured():
    """``measured("case", loss=..., grad=...)``: record the error a parity test actually measured (not just that it
    stayed under the tolerance).  Everything recorded is written to gpurun_out/parity_measured.json when the session ends."""
    def rec(name, **values):
        _MEASURED.setdefault(name, {}).update({k: (float(v) if isinstance(v, (int, float)) or hasattr(v, "__float__") else v)
                                               for k, v in values.items()})
    return rec


def pytest_sessionfinish(session, exitstatus):
    if not _MEASURED:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_measured.json")
        old = {}
        if os.path.exists(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_MEASURED)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)
    except OSError:
        pass
