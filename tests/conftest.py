import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The tests bind the in-tree shared library; build it when the sources are newer (nvcc
    cross-compiles sm_100a without a GPU).  The package itself never builds on import: a
    missing library is an error there."""
    from metacoag_b200 import build
    return build.build_library()
