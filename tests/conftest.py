import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native pieces exist (build() is what the driver runs anyway)."""
    from cable_b200 import lib
    from oracle import pyoracle
    if not (os.path.exists(lib.LIB_PATH) and os.path.exists(pyoracle.LIB_PATH) and os.path.exists(pyoracle.LIB_CR_PATH)):
        import __graft_entry__
        __graft_entry__.build()
