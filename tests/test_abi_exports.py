"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ofdm_b200.h declares; without a GPU the product path fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import PKG, ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "ofdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(PKG, "libofdm_b200.so")
    assert os.path.exists(lib_path), "run `python ofdm-based-systems_b200/build_native.py` (or __graft_entry__.build())"
    lib = ctypes.CDLL(lib_path)
    names = _declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ofdm_b200.h but not exported"
    lib.ofdm_b200_abi_version.restype = ctypes.c_int
    assert lib.ofdm_b200_abi_version() == 1


def test_binding_lists_the_same_symbols():
    from ofdm_based_systems import _native
    assert sorted(_native.EXPORTS) == _declared_functions()


def test_no_silent_cpu_fallback():
    import numpy as np
    from ofdm_based_systems import _native
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.Link(64, np.ones(1, complex), np.ones(64, complex), np.full(64, 4))
