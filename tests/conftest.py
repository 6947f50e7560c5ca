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
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))


@pytest.fixture(scope="session")
def built_lib():
    """libb200mel.so, (re)built in tree if the sources are newer."""
    from pytorch_sound_b200 import build

    build.build()
    from pytorch_sound_b200 import _lib

    return _lib
