import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle/libfmoracle.so), the reference harness when /root/reference is
    present, and the product library, once per session."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
