import glob
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
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def golden_instances():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "inst_*.npz")))


def load_instance(name):
    return np.load(os.path.join(GOLDEN, f"inst_{name}.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def dp_synth():
    return np.load(os.path.join(GOLDEN, "dp_synth.npz"))


@pytest.fixture(scope="session")
def sampler_kat():
    return np.load(os.path.join(GOLDEN, "sampler_kat.npz"))


CONTINUOUS = [n for n in golden_instances() if not n.endswith("_epi")]
EPISODIC = [n for n in golden_instances() if n.endswith("_epi")]
