"""pytest configuration.

  -m "not gpu"  oracle vs goldens / reference anchors, host logic, C-ABI symbol export (CPU only)
  -m gpu        parity tests proper: the CUDA path called through the C ABI vs the oracle
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")
    config.addinivalue_line("markers", "slow: larger sizes")


def golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def libpath():
    """The product library; built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from pharmsol_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from pharmsol_b200 import build
        build.build()
    return _lib.LIB_PATH


@pytest.fixture(scope="session")
def ps(libpath):
    import pharmsol_b200
    return pharmsol_b200


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
