import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference)")


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "golden_meta.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_arrays():
    return np.load(os.path.join(GOLDEN, "golden_arrays.npz"))


@pytest.fixture(scope="session")
def water():
    from qdk_chemistry_b200 import workloads as W
    return W.load_sparse_npz(os.path.join(GOLDEN, "h2o_ccpvdz.ints.npz"))


@pytest.fixture(scope="session")
def n2_18():
    from qdk_chemistry_b200 import workloads as W
    return W.load_sparse_npz(os.path.join(GOLDEN, "n2_14e18o.ints.npz"))


@pytest.fixture(scope="session")
def n2_6():
    from qdk_chemistry_b200 import workloads as W
    return W.load_sparse_npz(os.path.join(GOLDEN, "n2_6e6o.ints.npz"))
