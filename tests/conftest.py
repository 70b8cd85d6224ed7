import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """True when a CUDA device can be opened (asked of the driver through the product library itself)."""
    try:
        import ctypes
        import obvi_b200
        L = obvi_b200.lib()
        h = ctypes.c_void_p()
        if L.obvi_problem_create(0, ctypes.byref(h)) != 0:
            return False
        L.obvi_problem_destroy(h)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a machine without a CUDA device, so a plain `pytest` works on the CPU box."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device (the product has no CPU fallback; run these on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def ob():
    import obvi_b200
    return obvi_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.lib()
    return oracle_lib
