import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
        return cache[name]

    return load


@pytest.fixture
def qb_option():
    """Kernel-selection overrides of the library for one test (qb_set_option in include/qampy_b200.h); everything the
    test set is cleared afterwards."""
    from qampy_b200 import device

    touched = set()

    def set_option(name, value):
        touched.add(name)
        device.set_option(name, value)

    for name in ("TRAIN_KERNEL", "TRAIN_LPS", "BPS_KERNEL", "BPS_SPLIT"):   # nothing left over from the environment
        set_option(name, None)
    yield set_option
    for name in touched:
        device.set_option(name, None)
