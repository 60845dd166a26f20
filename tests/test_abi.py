"""CPU-side checks of the drop-in boundary: the C-ABI library exists in-tree, loads, and exports every
symbol include/pyrope_gpu.h declares (no compute calls: there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pyrope_gpu.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pyrope_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("pyrope_index_create", "pyrope_index_add_batch", "pyrope_index_search_batch",
                 "pyrope_index_build", "pyrope_index_delete_row", "pyrope_topk_merge_device",
                 "pyrope_last_error"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from pyrope_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"symbols declared in pyrope_gpu.h but not exported: {missing}"
    # and the Python binding names exactly the declared set
    assert sorted(_lib.SIGNATURES) == _declared()


def test_no_cpu_fallback_in_product():
    """The product package must not import or reference the oracle."""
    pkg = os.path.join(ROOT, "pyrope_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, \
                    f"{f} references the oracle"


def test_error_reporting_without_gpu():
    from pyrope_b200 import _lib
    L = _lib.load()
    assert L.pyrope_version() >= 100
    h = ctypes.c_void_p()
    # argument validation happens before any CUDA call, so it works on a CPU-only box
    rc = L.pyrope_index_create(_lib.FLAT, 0, _lib.L2, 0, 0, 0, ctypes.byref(h))
    assert rc == _lib.ERR_OUT_OF_RANGE and b"Dimension" in L.pyrope_last_error()
    rc = L.pyrope_index_create(_lib.IVF_PQ, 10, _lib.L2, 4, 3, 256, ctypes.byref(h))
    assert rc == _lib.ERR_INVALID_ARG and b"divisible" in L.pyrope_last_error()
    rc = L.pyrope_index_create(_lib.IVF_PQ, 16, _lib.L2, 4, 4, 257, ctypes.byref(h))
    assert rc == _lib.ERR_INVALID_ARG and b"256" in L.pyrope_last_error()
    rc = L.pyrope_index_create(7, 16, _lib.L2, 4, 4, 256, ctypes.byref(h))
    assert rc == _lib.ERR_INVALID_ARG
