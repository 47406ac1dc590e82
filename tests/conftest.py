import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "rfest_numpy_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference's own NumPy modules (oracle/make_golden.py)."""
    return np.load(GOLDEN)


@pytest.fixture(scope="session")
def reference_modules():
    """Live reference modules -- only in the build container."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    return ref_loader.load()


def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
