import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

_mod = None


def msfec_mod():
    """The ctypes binding mpi-msfec_b200/msfec_b200.py (directory name is not importable)."""
    global _mod
    if _mod is None:
        path = os.path.join(ROOT, "mpi-msfec_b200", "msfec_b200.py")
        spec = importlib.util.spec_from_file_location("msfec_b200", path)
        _mod = importlib.util.module_from_spec(spec)
        sys.modules["msfec_b200"] = _mod
        spec.loader.exec_module(_mod)
    return _mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def msfec():
    m = msfec_mod()
    if not os.path.exists(m.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return m
