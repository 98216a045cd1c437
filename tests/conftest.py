import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device: on a CPU-only box they are skipped (not failed) even without `-m "not gpu"`."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def diag_dir():
    d = ROOT / "gpurun_out" / "diag"
    d.mkdir(parents=True, exist_ok=True)
    return d


@pytest.fixture(scope="session")
def lib():
    from whisperjav_b200 import _lib
    return _lib.load()
