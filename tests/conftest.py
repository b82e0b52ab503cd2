"""Test configuration: `gpu` marker, library build, shared helpers."""

import os
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cap_lib():
    """The built C-ABI library (built on demand when nvcc is present)."""
    from openviic_b200 import build, cabi
    if not cabi.LIB_PATH.exists():
        build.build()
    return cabi.load_library()


@pytest.fixture(scope="session")
def device(cap_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
