"""ctypes binding of libmpinn_b200.so (the C ABI declared in include/mpinn_b200.h).

There is no CPU fallback: if the library is missing, or a call is made without a CUDA device,
the caller gets a RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPINN_B200_LIB selects a diagnostic build of the same library (scripts/build_variant.sh); it is never a fallback
LIB_PATH = os.environ.get("MPINN_B200_LIB") or os.path.join(_HERE, "csrc", "libmpinn_b200.so")

ABI_VERSION = 3
OP_LINEAR, OP_FOURIER = 0, 1
ACT_NONE, ACT_SINE, ACT_GELU = 0, 1, 2
FLAG_SAVE_SKIP, FLAG_ADD_SKIP, FLAG_TRAINABLE = 1, 2, 4
ENGINE_AUTO, ENGINE_FFMA, ENGINE_TCGEN05 = 0, 1, 2
ORDER_LAPLACE = 3
ORDER_HESSIAN = 4

# every symbol include/mpinn_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mpinn_abi_version": (c_int, []),
    "mpinn_last_error": (c_char_p, []),
    "mpinn_stash_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int64]),
    "mpinn_backward_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int64]),
    "mpinn_packed_bytes": (c_size_t, [c_void_p, c_int]),
    "mpinn_pack_params": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mpinn_jet_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mpinn_jet_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_size_t, c_int, c_void_p]),
    "mpinn_engine_supports": (c_int, [c_void_p, c_int, c_int, c_int, c_int]),
    "mpinn_residual_program": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "mpinn_adam_tick": (c_int, [c_void_p, c_float, c_float, c_void_p]),
    "mpinn_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float,
                                    c_float, c_float, c_float, c_void_p]),
    "mpinn_selftest_umma": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpinn_debug_counters": (c_int, [c_void_p, c_int]),
    "mpinn_sumsq": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "mpinn_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float,
                                c_float, c_float, c_float, c_float, c_void_p]),
}

_lib = None


class MpinnError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"multipinn_b200: CUDA library not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C multipinn_b200/csrc`. "
            "There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mpinn_abi_version() != ABI_VERSION:
        raise RuntimeError("multipinn_b200: libmpinn_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mpinn_last_error()
        raise MpinnError(f"{what} failed: {msg.decode() if msg else rc}")


def ops_array(program) -> np.ndarray:
    """int32 [n_ops, 8] C-contiguous array laid out exactly like ``mpinn_op_t[n_ops]``."""
    a = np.ascontiguousarray(np.asarray(program, dtype=np.int32).reshape(-1, 8))
    return a
