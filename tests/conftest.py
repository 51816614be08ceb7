import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import modulate_b200
        return modulate_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly rather than silently pass; but a plain
    # `pytest tests/` here (no marker filter) skips the GPU tests.
    if config.getoption("-m"):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "cycle_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mb():
    """The product package, initialised on cuda:0 (GPU tests only)."""
    import modulate_b200
    modulate_b200.init(0)
    return modulate_b200
