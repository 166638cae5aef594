import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def syn():
    import npcd_b200  # noqa: F401
    from npcd_b200 import synthetic

    return synthetic


@pytest.fixture(scope="session")
def weights(syn):
    return syn.make_weights(0)


@pytest.fixture(scope="session")
def cameras(syn):
    return syn.load_cameras()
