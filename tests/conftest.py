import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle  # test infrastructure: the CPU restatement of the reference
    return _oracle.get()


@pytest.fixture(scope="session")
def trn():
    """The product: ctypes view of the C-ABI library (trueno_b200/libtrueno_cuda.so)."""
    import trueno_b200
    return trueno_b200
