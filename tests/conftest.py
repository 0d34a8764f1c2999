import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_count():
    try:
        from pyblip_b200 import _lib
        return _lib.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    """Fails (does not skip) when no GPU is visible: -m gpu tests must run the CUDA path."""
    n = _gpu_count()
    assert n > 0, "no CUDA device visible: gpu-marked tests need a B200"
    return n
