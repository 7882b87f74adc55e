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
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def conf_from_golden(g, cls):
    """KLT configuration of a golden case (oracle/make_golden.py: conf_of)."""
    import ast
    base = dict(minDistance=10, blocksize=15, maxCorners=20000, matching_winsize=25,
                qualityLevel=0.1, xStart=0, tile_size=20000, laplacian_kernel_size=7,
                outliers_filtering=False, laplacian_invert_polarity=False)
    base.update(ast.literal_eval(str(g["conf_json"])))
    return cls(**base)
