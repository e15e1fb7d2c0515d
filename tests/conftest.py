import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (B200); run with -m gpu")


@pytest.fixture(scope="session")
def uav_lib():
    """Build (if stale) and load libuav_b200.so."""
    from upscale_a_video_b200 import build, _lib
    build.build()
    return _lib.load()
