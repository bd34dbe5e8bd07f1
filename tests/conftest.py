import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def _has_gpu() -> bool:
    try:
        from avlmaps_b200 import _lib

        return _lib.device_count() > 0
    except Exception:  # noqa: BLE001
        return False


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; GPU tests fail (not skip) if it is missing or no device is visible."""
    from avlmaps_b200 import _lib

    L = _lib.load()
    _lib.require_device()
    return L


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU SKIPS the gpu-marked tests (a CPU-only CI run stays green and real CPU
    regressions stay visible); an explicit `-m gpu` keeps them, so that they FAIL loudly there -- there is no CPU fallback
    to pass on silently."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="needs a B200: run with `-m gpu` under gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
