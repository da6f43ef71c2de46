import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_effb2():
    return dict(np.load(os.path.join(GOLDEN_DIR, "effb2_trm.npz")))


@pytest.fixture(scope="session")
def oracle_effb2(golden_effb2):
    """Oracle model with exactly the weights the golden vectors were produced with."""
    from oracle import caption_model as cm
    return cm.build_effb2_trm(int(golden_effb2["seed"]), bn_stats=golden_effb2["bn_stats"])


@pytest.fixture(scope="session")
def golden_wav(golden_effb2):
    from oracle import caption_model as cm
    g = golden_effb2
    wav, lens = cm.synth_wav(int(g["batch"]), int(g["n_samples"]), seed=int(g["wav_seed"]), ragged=True, varied=True)
    assert (lens.numpy() == g["wav_len"]).all()
    return wav, lens
