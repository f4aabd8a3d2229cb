"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/phertilizer_b200.h declares.
No compute is attempted here; without a GPU the build call must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "phertilizer_b200.h")).read()
    return re.findall(r"PH_API\s+[\w\s\*]+?\b(ph_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    from phertilizer_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 20 and set(syms) == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), s
    assert _lib.lib().ph_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from phertilizer_b200 import _lib
    from phertilizer_b200.store import ObservationStore

    with pytest.raises(_lib.PhError):
        ObservationStore.from_coo(4, 4, np.array([0, 1]), np.array([1, 2]), np.array([0, 1]), np.array([1, 1]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "phertilizer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
