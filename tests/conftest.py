import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the C-ABI library; GPU tests must run on the real kernels."""
    import torch
    from llmspeculativesampling_b200 import build, _cabi
    build.build()
    lib = _cabi.load()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return lib
