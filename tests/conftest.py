import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_ready():
    """A CUDA device and the built library: what every gpu-marked test needs."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as e:  # noqa: BLE001
        return False, f"torch unavailable: {e}"
    lib = os.path.join(ROOT, "ecseg_b200", "libecseg_b200.so")
    if not os.path.isfile(lib):
        return False, lib + " not built"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the gpu-marked tests instead of drowning the host-side
    results in Engine() errors.  On a GPU box nothing is skipped: a missing library there must fail loudly."""
    ok, why = _gpu_ready()
    if ok:
        return
    try:
        import torch
        if torch.cuda.is_available():
            return          # GPU present but library missing: let the tests fail, that is a broken build
    except Exception:  # noqa: BLE001
        pass
    skip = pytest.mark.skip(reason="needs a B200: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))

    return load
