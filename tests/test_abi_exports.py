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


def test_header_compiles_and_links_from_plain_c(tmp_path):
    """include/ofdm_b200.h is a C header: gcc -std=c99 -pedantic compiles a consumer, links it against the library and
    the program runs (with a GPU it simulates 128 000 bits; without one it checks the error path)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = tmp_path / "abi_check"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_check.c"), "-o", str(exe), "-L", PKG, "-lofdm_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
