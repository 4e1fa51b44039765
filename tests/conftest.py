"""pytest configuration: `gpu` marker (tests that need a B200), shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests must FAIL (not skip) when no device is present: there is no CPU fallback."""
    import torch

    assert torch.cuda.is_available(), "a CUDA device is required for -m gpu tests"
    from facemap_b200 import _lib

    _lib.require_device(0)
    return torch.device("cuda", 0)
