"""Test tiers (mirrors the reference's marker scheme, tests/conftest.py:66-88 there):

* unmarked tests run on CPU in the build container: the oracle against the golden vectors recorded
  from the reference, host-side logic, and the C-ABI surface of the built library;
* ``@pytest.mark.gpu`` tests are the parity tests proper: CUDA kernels (through the C ABI) against
  the oracle and the golden vectors. They are skipped when no CUDA device is present.
"""

from __future__ import annotations

import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent

for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(0)
    yield
