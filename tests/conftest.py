import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout), so a stuck collective cannot stall the suite")


def _gpu_count():
    try:
        from osiris_b200 import lib
        return lib.load().osgpu_device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a device must fail loudly, not skip silently; without -m the gpu
    # tests are skipped on CPU boxes so a bare `pytest tests` stays green here.
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    if _gpu_count() == 0:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """in-tree CUDA library (built by __graft_entry__.build / python -m osiris_b200.build)"""
    from osiris_b200 import build as b
    if not os.path.exists(b.OUT):
        b.build()
    return b.OUT
