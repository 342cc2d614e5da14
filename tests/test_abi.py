"""The C-ABI shared library: loads, exports every function include/mpres_b200.h declares, and fails
loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mpres_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpres_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(pkg):
    lib = pkg.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(pkg.EXPORTS) == names
    assert b"sm_100a" in lib.mpres_version()


def test_struct_layouts_match_reference_types(pkg):
    # src/types.cuh:46-49, 69-74, 85-104 on LP64
    assert ctypes.sizeof(pkg.mp_array_t) == 48 and ctypes.sizeof(pkg.mp_collection_t) == 32
    for N in (8, 16, 32, 64):
        dt = pkg.record_dtype(N)
        assert dt.itemsize == 4 * N + 40
        assert dt.fields["sign"][1] == 4 * N and dt.fields["exp"][1] == 4 * N + 4 and dt.fields["eval"][1] == 4 * N + 8


def test_no_cpu_fallback(pkg):
    """a constants-only context refuses every compute entry point; a device context cannot be made here"""
    import torch
    ctx = pkg.Context(8, -1)
    arr = pkg.mp_array_t()
    assert ctx.lib.mpres_array_init(ctx.h, ctypes.byref(arr), ctypes.c_size_t(4)) == -100
    assert ctx.lib.mpres_gemm(ctx.h, 111, 111, 1, 1, 1, ctypes.byref(arr), ctypes.byref(arr), 1, ctypes.byref(arr), 1,
                              ctypes.byref(arr), ctypes.byref(arr), 1, None, None) == -100
    assert ctx.lib.mpres_dot(ctx.h, 1, ctypes.byref(arr), 1, ctypes.byref(arr), 1, ctypes.byref(arr), None, None) == -100
    # the round-2 entry points too (norms, sparse product, conversions, division, solver, host GEMM, sharding)
    assert ctx.lib.mpres_asum(ctx.h, 1, ctypes.byref(arr), 1, ctypes.byref(arr), None) == -100
    assert ctx.lib.mpres_norm(ctx.h, 171, 1, ctypes.byref(arr), 1, ctypes.byref(arr), None) == -100
    assert ctx.lib.mpres_div(ctx.h, ctypes.byref(arr), ctypes.byref(arr), ctypes.byref(arr), None) == -100
    # the workspace budget is host state: settable and readable without a device
    assert ctx.workspace_bytes() == 0 and ctx.workspace_fallbacks() == 0
    ctx.set_workspace_limit(1 << 30)
    ctx.set_workspace_limit(0)
    ctx.close()
    if not torch.cuda.is_available():
        with pytest.raises(pkg.MpresError):
            pkg.Context(8, 0)


def test_product_does_not_import_oracle():
    """nothing under mpres-blas_b200/ may reference the oracle (SURVEY / task rule)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpres-blas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mpres_oracle" not in txt and "liboracle" not in txt, f
