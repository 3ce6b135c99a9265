import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The native library on a real sm_100 device; fails loudly (no skip) when -m gpu is requested
    on a box where it cannot run."""
    import torch
    from multimodal_sc_b200 import _lib
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    _lib.require_device()
    return _lib.load()
