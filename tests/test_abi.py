"""CPU: the C-ABI library builds/loads and exports every symbol include/sc_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = _declared_symbols()
    assert "sc_chamfer_forward" in syms and "sc_abi_version" in syms


def test_library_exports_every_declared_symbol():
    from shapeclipper_b200 import build
    so = build.build()
    lib = ctypes.CDLL(so)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.sc_abi_version.restype = ctypes.c_int
    from shapeclipper_b200 import _lib
    assert lib.sc_abi_version() == _lib.ABI_VERSION


def test_loader_declares_every_symbol():
    from shapeclipper_b200 import _lib
    L = _lib.lib()
    for s in _declared_symbols():
        fn = getattr(L, s)
        assert fn.restype is not None or s.endswith("_bytes")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "shapeclipper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f == "build.py", f
