import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    from helpers import load_golden

    return load_golden()


@pytest.fixture(scope="session")
def tool():
    """One GPU solver handle for the whole session.  Fails loudly (no skip, no fallback) if the CUDA library or
    the device is missing: a GPU test must never pass on anything but the CUDA path."""
    from cs_pathplan_b200 import TrajectoryGeneratorTool

    t = TrajectoryGeneratorTool(0)
    yield t
    t.close()
