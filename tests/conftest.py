import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def grb():
    """The product package with its CUDA library built in-tree (nvcc cross-compiles without a GPU)."""
    from graniitti_b200 import build
    build.build_lib()
    import graniitti_b200
    return graniitti_b200


@pytest.fixture(scope="session")
def msa():
    import numpy as np
    from helpers import GOLDEN
    return np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
