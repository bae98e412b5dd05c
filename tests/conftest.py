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


TRAINED_CASES = ["t_photo_64x96", "t_photo_53x77"]     # weights trained by the reference's own mode: train (ckpt_A_trained.npz)
TRAIN_CASES = ["train_a_2x32x32", "train_b_3x24x40", "train_t_2x64x32"]     # one backward pass of the reference's training step (make_golden.py)
FORWARD_CASES = ["fwd_a_photo_64x96", "fwd_a_noise_32x64", "fwd_a_checker_32x32", "fwd_b_photo_64x96", "fwd_b_photo_36x52"]


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def oracle_config_for(name):
    from oracle import llicti_oracle as O
    return O.OracleConfig() if name.startswith(("a_", "fwd_a_", "t_", "train_a_", "train_t_")) else O.OracleConfig(dwtlevels=(0, 1), chs=60)


def trained_state_dict():
    """llicti_A weights short-trained by the unmodified reference's `mode: train` on synthetic patches
    (tools/train_reference_ckpt.py); the reference's shipped checkpoint is absent."""
    import numpy as np
    with np.load(os.path.join(GOLDEN, "ckpt_A_trained.npz")) as z:
        return {k: z[k] for k in z.files}


def train_state_dict(name):
    """Weights of a training-step fixture: the reference-trained checkpoint for train_t_*, else generic (jittered) ones."""
    from oracle import llicti_oracle as O
    if name.startswith("train_t_"):
        return trained_state_dict()
    return O.jittered_state_dict(oracle_config_for(name), seed=1337)
