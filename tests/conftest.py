import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libtqsim.so built in-tree (nvcc cross-compiles without a GPU)."""
    from tensorrl_qas_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle
