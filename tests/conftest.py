import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_DIR = "/root/reference"  # only present in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def reference_module():
    """The live reference, for tests that can only run in the build container."""
    if not os.path.exists(os.path.join(REFERENCE_DIR, "train.py")):
        pytest.skip("reference not present (GPU box)")
    sys.path.insert(0, REFERENCE_DIR)
    import train
    return train
