import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the oracle's C coder."""
    import __graft_entry__ as ge
    ge.build()
    from llicti_b200 import _lib
    return _lib.load()


GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["a_photo_33x47", "a_photo_53x77", "a_photo_64x96", "a_const_40x72", "a_noise_32x64",
                "a_checker_35x32", "b_photo_64x96", "b_photo_37x53"]


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def oracle_config_for(name):
    from oracle import llicti_oracle as O
    return O.OracleConfig() if name.startswith("a_") else O.OracleConfig(dwtlevels=(0, 1), chs=60)
