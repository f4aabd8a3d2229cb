import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_usable() -> str:
    """'' if a CUDA device can be used through the C-ABI library, else the reason."""
    import ctypes as C

    try:
        from phertilizer_b200 import _lib

        n = C.c_int(0)
        rc = _lib.lib().ph_device_count(C.byref(n))
        if rc != 0:
            return _lib.lib().ph_last_error().decode()
        return "" if n.value > 0 else "no CUDA device"
    except Exception:  # library missing / not loadable: do NOT skip -- the GPU tests must fail loudly
        return ""


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (one probe per session) on a box without a usable CUDA device, so that a plain `pytest tests`
    there reports the CPU-side signal instead of one CUDA error per GPU test.  `-m gpu` on a GPU box is unaffected."""
    if not any("gpu" in it.keywords for it in items):
        return
    why = _gpu_usable()
    if why:
        skip = pytest.mark.skip(reason=f"no usable GPU: {why}")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def example_golden():
    from tests.golden_util import load_golden
    return load_golden("example")


@pytest.fixture(scope="session")
def synth_golden():
    from tests.golden_util import load_golden
    return load_golden("synth_t0")
