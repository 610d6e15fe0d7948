import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def eval_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "eval_golden.npz"))


@pytest.fixture(scope="session")
def net_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "net_golden.npz"))


@pytest.fixture(scope="session")
def train_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "train_golden.npz"))
