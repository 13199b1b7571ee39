import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without CUDA skips the gpu-marked tests instead of failing at the first one
    (the product has no CPU fallback: its entry points raise GcmfError there, which is tested separately)."""
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def reference_goldens():
    """The reference's own 18 golden arrays (f4), decoded by tests/golden/make_golden.py."""
    return dict(np.load(os.path.join(GOLDEN_DIR, "reference_goldens.npz")))


@pytest.fixture(scope="session")
def ref_outputs():
    """Outputs of the live reference captured by tests/golden/make_golden.py (fp64)."""
    return dict(np.load(os.path.join(GOLDEN_DIR, "ref_outputs.npz")))


def rel_l2(a, b):
    """relative L2 error over the finite entries of b; NaN patterns must agree exactly."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN masks differ"
    ok = ~np.isnan(b)
    den = np.linalg.norm(b[ok])
    num = np.linalg.norm(a[ok] - b[ok])
    return num / den if den > 0 else num
