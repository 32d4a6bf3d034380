"""ctypes binding of ``libfa2_sm100.so`` (C ABI in ``include/fa2_sm100.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller gets an
exception. The reference fails the same way on a machine without its Triton backend
(``assert q.is_cuda`` at src/forward/caller.py:41).
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_NAME = "libfa2_sm100.so"
LIB_PATH = Path(os.environ.get("FA2_SM100_LIB", _PKG_DIR / LIB_NAME))

# error codes of include/fa2_sm100.h
E_UNSUPPORTED_HEAD_DIM = -1
E_UNSUPPORTED_DTYPE = -2
E_BAD_SHAPE = -3
E_BAD_ALIGNMENT = -4
E_NULL_POINTER = -5
E_WORKSPACE_TOO_SMALL = -6
E_NO_DRIVER = -7

# the `deterministic` argument of fa2_sm100_bwd
DQ_ATOMIC, DQ_DETERMINISTIC, DQ_FUSED_ORDERED, DQ_NONE, DQ_ONLY, DQ_DET_SPILL, DQ_SPILL_DKDV, DQ_SPILL_DQ = range(8)

DTYPE_FP16 = 16
DTYPE_BF16 = 17
DTYPE_FP32 = 32

_I64x3 = ctypes.c_int64 * 3
_c_int, _c_float, _c_void_p, _c_size_t, _c_int64 = (
    ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "fa2_sm100_version": (_c_int, []),
    "fa2_sm100_last_error": (ctypes.c_char_p, []),
    "fa2_sm100_fwd": (_c_int, [_c_void_p] * 5 + [_c_int] * 6 + [ctypes.POINTER(_c_int64)] * 4
                      + [_c_int64, _c_float, _c_int, _c_int, _c_void_p]),
    "fa2_sm100_fwd_ext": (_c_int, [_c_void_p] * 5 + [_c_int] * 6 + [ctypes.POINTER(_c_int64)] * 4
                          + [_c_int64, _c_float, _c_int, _c_int]
                          + [_c_void_p, _c_int, ctypes.POINTER(_c_int64), _c_float, ctypes.c_uint64, ctypes.c_uint32, _c_void_p]),
    "fa2_sm100_varlen_workspace_bytes": (_c_size_t, [_c_int]),
    "fa2_sm100_fwd_varlen": (_c_int, [_c_void_p] * 5 + [_c_int] * 5 + [ctypes.POINTER(_c_int64)] * 4
                             + [_c_int64, _c_float, _c_int, _c_int] + [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "fa2_sm100_bwd_delta_varlen": (_c_int, [_c_void_p] * 3 + [_c_int] * 4 + [ctypes.POINTER(_c_int64)] * 2
                                   + [_c_int64, _c_int, _c_void_p, _c_void_p]),
    "fa2_sm100_bwd_varlen": (_c_int, [_c_void_p] * 10 + [_c_size_t] + [_c_int] * 5 + [ctypes.POINTER(_c_int64)] * 7
                             + [_c_int64, _c_float, _c_int, _c_int, _c_int, _c_int]
                             + [_c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "fa2_sm100_bwd_delta": (_c_int, [_c_void_p] * 3 + [_c_int] * 4 + [ctypes.POINTER(_c_int64)] * 2
                            + [_c_int64, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "fa2_sm100_bwd_workspace_bytes": (_c_size_t, [_c_int] * 7),
    "fa2_sm100_bwd_workspace_bytes_gqa": (_c_size_t, [_c_int] * 8),
    "fa2_sm100_bwd": (_c_int, [_c_void_p] * 10 + [_c_size_t] + [_c_int] * 6 + [ctypes.POINTER(_c_int64)] * 7
                      + [_c_int64, _c_float, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "fa2_sm100_bwd_ext": (_c_int, [_c_void_p] * 10 + [_c_size_t] + [_c_int] * 6 + [ctypes.POINTER(_c_int64)] * 7
                          + [_c_int64, _c_float, _c_int, _c_int, _c_int, _c_int]
                          + [_c_void_p, _c_int, ctypes.POINTER(_c_int64), _c_void_p]),
    "fa2_sm100_dq_convert": (_c_int, [_c_void_p] * 2 + [_c_int] * 4 + [ctypes.POINTER(_c_int64)]
                             + [_c_float, _c_int, _c_void_p]),
    "fa2_sm100_copy2d_async": (_c_int, [_c_void_p, _c_size_t, _c_void_p, _c_size_t, _c_size_t, _c_size_t, _c_int, _c_void_p]),
}

_lib = None
_lock = threading.Lock()


class Fa2LibraryError(RuntimeError):
    """libfa2_sm100.so is missing, stale or returned a CUDA error."""


def load():
    """Load (once) and return the shared library with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise Fa2LibraryError(
                f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {_PKG_DIR / 'csrc'}`. There is no CPU or PyTorch fallback for this path.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means a stale build
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def strides3(t):
    """(batch, head, seq) element strides of a [B, S, H, D] tensor - the triple the reference passes to its
    kernels as ``x.stride(0), x.stride(2), x.stride(1)`` (src/forward/caller.py:93-96)."""
    return _I64x3(t.stride(0), t.stride(2), t.stride(1))


def bias_args(bias):
    """(pointer, dtype code, (batch, head, row) element strides) of an attention bias [1 | B, 1 | Hq, Sq, Sk]; size-1
    dimensions broadcast with stride 0 (what src/utils.py:57-77 computes). ``None`` -> (None, 0, None)."""
    if bias is None:
        return None, 0, None
    import torch
    code = {torch.float16: DTYPE_FP16, torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_FP32}[bias.dtype]
    sb = bias.stride(0) if bias.size(0) > 1 else 0
    sh = bias.stride(1) if bias.size(1) > 1 else 0
    return bias.data_ptr(), code, _I64x3(sb, sh, bias.stride(2))


def check(code: int, what: str):
    """Turn a non-zero return code into the exception type the reference would have raised."""
    if code == 0:
        return
    msg = load().fa2_sm100_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg} (code {code})"
    if code in (E_UNSUPPORTED_HEAD_DIM, E_UNSUPPORTED_DTYPE):
        raise NotImplementedError(text)
    if code < 0:
        raise ValueError(text)
    raise Fa2LibraryError(text)
