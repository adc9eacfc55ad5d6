import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

TRAJECTORIES = ["c1_seed1", "c1_seed42", "c2_abrupt", "c2_const_k46", "c1_linear", "c1_log", "c1_abrupt_stop", "c1_ka1",
                "big_blocks", "isolated"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def pkg():
    import importlib
    return importlib.import_module("bipartitesbm-mcmc_b200")


@pytest.fixture(scope="session")
def host(pkg):
    return pkg.host
