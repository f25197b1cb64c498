import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import pdynamo_mirror_b200
    return pdynamo_mirror_b200


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % name), allow_pickle=False)
