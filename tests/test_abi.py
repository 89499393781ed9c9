"""C-ABI surface: the library loads, exports every symbol include/xsb200.h declares, and fails loudly
(no CPU fallback) when no GPU is usable.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import exastamp_b200 as xsb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "xsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xsb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    xsb.build()
    L = xsb.load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libxsb200.so does not export %s" % s
    assert sorted(xsb.ABI_SYMBOLS) == syms, "python binding list out of sync with include/xsb200.h"


def test_version_string():
    assert b"sm_100a" in xsb.load_library().xsb_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(xsb.XsbError) as e:
        xsb.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_null_context_is_rejected():
    L = xsb.load_library()
    assert L.xsb_sync(None) != 0
    assert L.xsb_grid_set(None, None) != 0
    assert L.xsb_pair_force(None, 0, None, 0, 1.0, 0) != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "exastamp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, "%s mentions the oracle" % f


def test_every_abi_symbol_has_a_typed_ctypes_binding():
    """the ctypes mirror declares argument (or result) types for every entry point: an untyped call would pass
    doubles and 64-bit sizes through the default int conversion"""
    import re
    src = open(os.path.join(ROOT, "exastamp_b200", "__init__.py")).read()
    for s in xsb.ABI_SYMBOLS:
        assert ("L.%s.argtypes" % s) in src or ("L.%s.restype" % s) in src, "%s is not typed" % s
