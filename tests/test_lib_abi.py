"""The C-ABI library loads without a GPU and exports every symbol include/mpinn_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from multipinn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mpinn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpinn_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_all_symbols():
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.mpinn_abi_version() == _lib.ABI_VERSION


def test_size_queries_without_gpu():
    """Pure host arithmetic: stash / workspace / packed sizes (no compute calls)."""
    lib = _lib.load()
    ops = _lib.ops_array([[0, 2, 64, 0, 128, 1, 0, 0], [0, 64, 64, 192, 4288, 2, 0, 0],
                          [0, 64, 1, 4352, 4416, 0, 0, 0]])
    p = ops.ctypes.data_as(ctypes.c_void_p)
    # J=5 jets, two stashed layers of width 64, points padded to 128
    assert lib.mpinn_stash_bytes(p, 3, 2, 2, 1000) == 5 * 64 * 2 * 1024 * 4
    assert lib.mpinn_stash_bytes(p, 3, 2, 0, 1000) == 1 * 64 * 2 * 1024 * 4
    assert lib.mpinn_backward_workspace_bytes(p, 3, 4417) >= 148 * 4417 * 4
    assert lib.mpinn_packed_bytes(p, 3) >= (2 * 64 + 64 * 8 + 64 * 64 * 2 + 64 * 8 + 64) * 4
    bad = _lib.ops_array([[0, 2, 64, 0, 128, 1, 0, 0], [0, 32, 64, 192, 4288, 2, 0, 0]])
    assert lib.mpinn_stash_bytes(bad.ctypes.data_as(ctypes.c_void_p), 2, 2, 2, 10) == 0


def test_no_cpu_fallback():
    """Without CUDA the product path must fail loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import multipinn_b200 as mp
    from multipinn_b200 import engine

    old = engine._set_backend_for_tests(None)
    try:
        model = mp.FNN(2, 1, [8, 8])
        with pytest.raises(RuntimeError):
            model(torch.zeros(4, 2))
    finally:
        engine._set_backend_for_tests(old)
