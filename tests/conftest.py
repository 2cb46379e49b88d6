import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding

    return oracle_binding.load()


@pytest.fixture(scope="session")
def handle():
    """One C-ABI handle bound to torch's current stream, like the reference's BLAS_Test fixture
    (reference tests/common.h:9-24)."""
    import torch
    import rtatblas_b200 as rt

    assert torch.cuda.is_available(), "gpu tests need a GPU"
    h = rt.Handle(stream=torch.cuda.current_stream())
    yield h
    torch.cuda.synchronize()
    h.close()
