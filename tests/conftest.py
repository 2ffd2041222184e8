import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference EMD extension (oracle/_ref/emd.so), or None when it was not built."""
    from oracle import build_ref
    return build_ref.load_ref()
