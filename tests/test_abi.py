"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares, and
rejects bad arguments without touching a GPU; the Python surface mirrors the reference's."""
import ctypes
import inspect
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "fa2_sm100.h"


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(fa2_sm100_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = header_functions()
    for required in ("fa2_sm100_fwd", "fa2_sm100_bwd_delta", "fa2_sm100_bwd_workspace_bytes", "fa2_sm100_bwd",
                     "fa2_sm100_dq_convert", "fa2_sm100_last_error", "fa2_sm100_version"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from fa2_triton_b200 import _lib
    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/fa2_sm100.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.fa2_sm100_version() == 13


def test_workspace_size_formula():
    from fa2_triton_b200 import _lib
    lib = _lib.load()
    # atomic / fused-ordered: fp32 accumulator over Sq padded to 128 + one ticket per (b, h, tile) + the counter that
    # numbers the CTAs of the ticket-ordered kernel in the order they start, 16-byte rounded
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 1, 0) == 4 * 32 * 4096 * 128 * 4 + 4 * 32 * 32 * 4 + 16
    assert lib.fa2_sm100_bwd_workspace_bytes(1, 1, 1000, 1000, 64, 0, 2) == 1024 * 64 * 4 + 48
    assert lib.fa2_sm100_bwd_workspace_bytes(0, 1, 1, 1, 64, 0, 0) == 0
    # deterministic / dQ only: dQ has its own kernel, nothing to accumulate into - just its scheduler's counter
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 1, 1) == 256
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 1, 4) == 256
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 1, 3) == 256
    # deterministic with spilled dS: 1 KB + one 32 KB tile per visible 128 x 128 block pair
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 1, 5) == 1024 + 4 * 32 * (32 * 33 // 2) * 32768
    assert lib.fa2_sm100_bwd_workspace_bytes(4, 32, 4096, 4096, 128, 0, 5) == 1024 + 4 * 32 * 32 * 32 * 32768
    # bottom-right aligned causal mask, Sq = 300 (3 query blocks), Sk = 1000: the blocks see 7, 8 and 8 key blocks
    assert lib.fa2_sm100_bwd_workspace_bytes(1, 1, 300, 1000, 64, 1, 5) == 1024 + (7 + 8 + 8) * 32768


def test_argument_errors_need_no_gpu():
    from fa2_triton_b200 import _lib
    lib = _lib.load()
    s = (ctypes.c_int64 * 3)(128 * 8 * 64, 64, 8 * 64)
    buf = ctypes.create_string_buffer(64)
    ptr = (ctypes.addressof(buf) + 15) & ~15
    rc = lib.fa2_sm100_fwd(ptr, ptr, ptr, ptr, ptr, 1, 8, 8, 128, 128, 44, s, s, s, s, 128, 0.1, 0, 17, None)
    assert rc == _lib.E_UNSUPPORTED_HEAD_DIM and b"head_dim 44" in lib.fa2_sm100_last_error()
    rc = lib.fa2_sm100_fwd(ptr, ptr, ptr, ptr, ptr, 1, 8, 8, 128, 128, 264, s, s, s, s, 128, 0.1, 0, 17, None)
    assert rc == _lib.E_UNSUPPORTED_HEAD_DIM
    # an expanded tensor (zero stride over a dimension of extent > 1) is refused, not silently mis-addressed
    z = (ctypes.c_int64 * 3)(128 * 8 * 64, 0, 8 * 64)
    rc = lib.fa2_sm100_fwd(ptr, ptr, ptr, ptr, ptr, 1, 8, 8, 128, 128, 64, z, s, s, s, 128, 0.1, 0, 17, None)
    assert rc == _lib.E_BAD_SHAPE and b"expanded" in lib.fa2_sm100_last_error()
    rc = lib.fa2_sm100_fwd(ptr, ptr, ptr, ptr, ptr, 1, 8, 3, 128, 128, 64, s, s, s, s, 128, 0.1, 0, 17, None)
    assert rc == _lib.E_BAD_SHAPE
    rc = lib.fa2_sm100_fwd(ptr, ptr, ptr, ptr, ptr, 1, 8, 8, 128, 128, 64, s, s, s, s, 128, 0.1, 0, 32, None)
    assert rc == _lib.E_UNSUPPORTED_DTYPE
    with pytest.raises(NotImplementedError):
        _lib.check(rc, "fwd")
    with pytest.raises(ValueError):
        _lib.check(_lib.E_BAD_SHAPE, "fwd")


def test_python_surface_matches_reference_signatures():
    import fa2_triton_b200 as pkg
    want = ["q", "k", "v", "attention_mask", "attention_bias", "dropout_p", "causal", "softmax_scale", "dropout_seed"]
    assert list(inspect.signature(pkg.flash_attn_func).parameters) == want  # src/wrapper.py:89-99
    assert list(inspect.signature(pkg.FlashAttnFunc.forward).parameters) == ["ctx"] + want  # src/wrapper.py:13-24
    fwd = list(inspect.signature(pkg._flash_attn_forward).parameters)
    # the reference's nine parameters in its order (src/forward/caller.py:12-22); `out` and `dropout_offset` are trailing
    # extensions
    assert fwd[:9] == ["q", "k", "v", "attention_mask", "bias", "dropout_p", "causal", "softmax_scale", "dropout_seed"]
    assert fwd[9:] == ["out", "dropout_offset"]
    bwd = list(inspect.signature(pkg._flash_attn_backward).parameters)
    assert bwd[:12] == ["dO", "q", "k", "v", "bias", "attention_mask", "o", "lse", "dropout_p", "causal",
                        "softmax_scale", "dropout_seed"]  # src/backward/caller.py:14-27


def test_cpu_tensors_are_rejected_like_the_reference():
    import fa2_triton_b200 as pkg
    q = torch.zeros(1, 128, 2, 64, dtype=torch.bfloat16)
    with pytest.raises(AssertionError):  # src/forward/caller.py:41 `assert q.is_cuda ...`
        pkg.flash_attn_func(q, q, q)
    with pytest.raises(AssertionError):  # src/forward/caller.py:40 dtype check
        pkg.flash_attn_func(q.float(), q.float(), q.float())
    with pytest.raises(AssertionError):  # dropout does not change that
        pkg.flash_attn_func(q, q, q, dropout_p=0.1)
    with pytest.raises(AssertionError):  # src/utils.py:81-82
        pkg.flash_attn_func(q, q, q, dropout_p=1.0)
    with pytest.raises(NotImplementedError):  # src/utils.py:87-88: no dropout in the backward pass
        pkg._flash_attn_backward(q, q, q, q, None, None, q, torch.zeros(1, 2, 128), 0.1, False, None, 1)


def test_product_package_does_not_import_the_oracle():
    for path in (ROOT / "fa2_triton_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{path} imports oracle/"
