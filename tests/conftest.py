import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_scripts():
    """Known-answer data shipped by the reference (check.py, check_debug.py, py_debug.txt)."""
    return np.load(os.path.join(GOLDEN, "reference_scripts.npz"))


@pytest.fixture(scope="session")
def golden_ref():
    """Outputs of the unmodified reference classes on small grids (tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def golden_msg():
    """More MSGSolver runs of the unmodified reference: exact-error rule, no true solution, 64 x 64, iteration cap."""
    return np.load(os.path.join(GOLDEN, "reference_outputs_msg.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.Oracle.lib()
    return oracle
