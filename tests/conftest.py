import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_oracle, load_package  # noqa: E402

load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    return load_oracle()


@pytest.fixture(scope="session")
def cuda_lib():
    """libpdenlp_cuda.so, built on demand (nvcc cross-compiles without a GPU)."""
    from pdenlp_b200 import build, capi
    build.build_cuda()
    return capi.load()
