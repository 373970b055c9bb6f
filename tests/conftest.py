import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diffusion-integer-programming_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand here (nvcc cross-compiles without a GPU)."""
    from dip_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH):
        build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)
