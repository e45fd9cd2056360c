import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable():
    """reason string when the CUDA path cannot run here (no device / library not built), else None"""
    try:
        import ctypes
        from picaso_b200 import _lib
        lib = _lib.load_library()
        n = ctypes.c_int(0)
        lib.pb_device_count(ctypes.byref(n))
        return None if n.value > 0 else "no CUDA device visible"
    except Exception as e:  # library missing / not loadable
        return "libpicaso_b200.so not loadable: %r" % (e,)


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a CPU box skips the gpu-marked tests instead of erroring in every one of them
    (`-m gpu` on a box without a device still reports them as skipped, not passed)"""
    if not any("gpu" in it.keywords for it in items):
        return
    why = _gpu_unavailable()
    if why is None:
        return
    skip = pytest.mark.skip(reason=why)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
