import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:  # test modules share helpers (test_cpp_host_gpu.run / write_wav)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def section71():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "section71.npz")))


@pytest.fixture(scope="session")
def sample_excerpt():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "sample_excerpt.npz")))


@pytest.fixture(scope="session")
def synthetic_small():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "synthetic_small.npz")))
