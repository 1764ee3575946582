import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.setdefault("E2E_CTC_TEST_ENV", "1")   # tests toggle the kernel-plan switches between calls (engine._plan_env)
os.environ.setdefault("OMP_NUM_THREADS", "8")  # unset, tiny torch CPU ops stall ~50 ms in this image


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Make sure libe2e_ctc.so exists (nvcc cross-compiles without a GPU)."""
    from end2end_b200 import build
    return build.build()
