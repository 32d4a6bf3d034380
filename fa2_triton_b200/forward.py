"""Host caller of the forward kernel - the restatement of ``_flash_attn_forward``
(reference src/forward/caller.py:12-122): validate, allocate O and LSE, launch."""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def encode_dtype(x: Tensor) -> int:
    """dtype -> the reference's integer code (src/utils.py:102-109)."""
    if x.dtype == torch.float16:
        return _lib.DTYPE_FP16
    if x.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    if x.dtype == torch.float32:
        return _lib.DTYPE_FP32
    raise ValueError(x.dtype)


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


SUPPORTED_HEAD_DIMS = (64, 128, 256)  # the kernel instantiations
FUSED_HEAD_DIMS = (64, 128)            # ... of the fused (atomic / ticket-ordered) backward


def runs_natively(head_dim: int) -> bool:
    """Head dims that go to the kernels as they are: multiples of 8 up to 256. TMA zero-fills the columns between the
    tensor's head dim and the kernel's (64, 128 or 256) on load and clips them on store - no padded copies."""
    return 8 <= head_dim <= SUPPORTED_HEAD_DIMS[-1] and head_dim % 8 == 0


def padded_head_dim(head_dim: int, dims=None) -> int:
    """Any other head dim up to 256 (odd ones like 59, 111 or 207: TMA needs 16-byte strides) runs on zero-padded copies:
    zero columns add nothing to Q K^T and produce zero columns of O, dQ, dK, dV. The reference pads to the next power of
    two inside its kernels the same way (src/forward/caller.py:77-78, PADDED_HEADS). Without ``dims`` the copy is padded
    to the next multiple of 8 (what TMA needs); with ``dims`` to the first of those head dims that holds it."""
    if dims is None:
        d = max(8, (head_dim + 7) // 8 * 8)
        if d <= SUPPORTED_HEAD_DIMS[-1]:
            return d
    else:
        for d in dims:
            if head_dim <= d:
                return d
    raise NotImplementedError(f"head_dim {head_dim}: this build supports head dims up to {SUPPORTED_HEAD_DIMS[-1]}")


def pad_last(x: Tensor, d: int) -> Tensor:
    return x if x.shape[-1] == d else torch.nn.functional.pad(x, (0, d - x.shape[-1]))


def kernel_addressable(x: Tensor) -> Tensor:
    """``x`` itself if TMA can address it, else a contiguous copy.

    The reference's Triton kernels take any strides (pointer arithmetic, src/forward/caller.py:93-96). TMA wants a
    contiguous last dimension, a 16-byte aligned base, every other stride a multiple of 8 elements - and no zero stride
    over a dimension of extent > 1: an EXPANDED tensor (k / v expanded over heads, or the dO autograd produces for
    ``out.sum()``) would make the rows alias. Such tensors are materialised here, like ``src/wrapper.py:41-43`` does for
    a non-contiguous last dimension."""
    if x.dim() != 4:
        return x
    ok = x.stride(-1) == 1 and x.data_ptr() % 16 == 0
    for dim in range(3):
        st = x.stride(dim)
        if x.size(dim) > 1 and (st == 0 or st % 8 != 0):
            ok = False
    return x if ok else x.contiguous()


def device_sequence_lengths(attention_mask: Tensor) -> Tensor:
    """int32 [B] lengths of a right-padded [B, S] mask, computed and kept on the device."""
    return attention_mask.sum(dim=1, dtype=torch.int32).contiguous()


def varlen_in_kernel(q: Tensor, k: Tensor, v: Tensor) -> bool:
    """Variable-length batches run in one launch when the kernels can address the tensors as they are (head dim a
    multiple of 8 up to 128, 16-bit dtype); anything else takes the per-sequence path below."""
    return (runs_natively(q.shape[-1]) and q.dtype in (torch.float16, torch.bfloat16)
            and os.environ.get("FA2_VARLEN_PER_SEQUENCE", "0") != "1")


def sequence_lengths(attention_mask: Tensor):
    """Per-sequence lengths of a right-padded [B, S] mask (what src/utils.py:14 reads with .item(); one host sync)."""
    return [int(n) for n in attention_mask.sum(dim=1).tolist()]


def check_bias(bias: Optional[Tensor], q: Tensor, batch: int, nheads_q: int, seqlen_q: int, seqlen_k: int):
    """Validate an attention bias ``[1 | B, 1 | Hq, Sq, Sk]`` and return it with a contiguous last dimension.

    Same acceptance rule and exception types as the reference's ``infer_bias_strides`` (src/utils.py:57-77) - with the
    head dimension compared against the head COUNT (the reference compares a stride, :70, which rejects per-head
    biases unless the stride happens to equal the count). The bias is indexed by the query head in both passes
    (the reference's forward uses the kv head, src/forward/kernel.py:141, its backward the query head,
    src/backward/kernel.py:143; they coincide for its own tests, which use a ``[1, 1, Sq, Sk]`` bias)."""
    if bias is None:
        return None
    assert bias.is_cuda
    assert bias.dim() == 4 and (bias.size(2) == seqlen_q and bias.size(3) == seqlen_k), f"{bias.shape = }"
    if bias.size(0) not in (1, batch):
        raise ValueError(f"Attention bias has {bias.size(0) = } while {batch = }")
    if bias.size(1) not in (1, nheads_q):
        raise ValueError(f"Attention bias has {bias.size(1) = } while {nheads_q = }")
    if bias.dtype not in (q.dtype, torch.float32):
        raise ValueError(f"Attention bias has dtype {bias.dtype}: expected {q.dtype} or torch.float32")
    return bias if bias.stride(-1) == 1 else bias.contiguous()


def handle_dropout(dropout_p: float, dropout_seed: Optional[int], is_forward: bool) -> int:
    """The reference's rule (src/utils.py:80-88): no dropout -> seed 0; forward draws a seed when none is given; the
    backward pass has no dropout."""
    assert dropout_p >= 0, f"Dropout probability {dropout_p = } must be above 0."
    assert dropout_p < 1, f"Dropout probability {dropout_p = } must be strictly below 1."
    if dropout_p == 0:
        return 0
    if is_forward:
        return torch.randint(low=0, high=2**32, size=(1,)).item() if dropout_seed is None else dropout_seed
    raise NotImplementedError("Backward pass does not yet support dropout.")


def _flash_attn_forward(
    q: Tensor,  # [batch_size, seqlen_q, num_heads_q, head_dim]
    k: Tensor,  # [batch_size, seqlen_k, num_heads_kv, head_dim]
    v: Tensor,  # [batch_size, seqlen_k, num_heads_kv, head_dim]
    attention_mask: Optional[Tensor],
    bias: Optional[Tensor],
    dropout_p: float = 0.0,
    causal: bool = False,
    softmax_scale: Optional[float] = None,
    dropout_seed: Optional[int] = None,
    out: Optional[Tensor] = None,
    dropout_offset: int = 0,
) -> Tuple[Tensor, Tensor, float, int]:
    """Returns ``(o, lse, softmax_scale, dropout_seed)`` like src/forward/caller.py:122.
    ``out`` (extension): write O into this tensor (any strides with a contiguous last dimension).
    ``dropout_offset`` (extension): added to every dropout counter modulo 2**32 (a sequence of a packed batch).

    ``lse`` is fp32 ``[B, Hq, ceil(Sq/128)*128]`` in the log2 domain (src/forward/caller.py:73-74,
    src/forward/kernel.py:119); entries past ``Sq`` are 0.
    """
    batch, seqlen_q, nheads_q, head_dim = q.shape
    _, seqlen_k, nheads_kv, _ = k.shape
    expected_kv_shape = (batch, seqlen_k, nheads_kv, head_dim)
    if q.is_cuda and k.is_cuda and v.is_cuda and runs_natively(head_dim):
        q, k, v = kernel_addressable(q), kernel_addressable(k), kernel_addressable(v)
    if attention_mask is not None:
        assert bias is None, "Attention mask is not supported along with attention bias. Just use bias instead."
        with_dropout = bool(dropout_p and dropout_p > 0)
        if with_dropout:
            dropout_seed = handle_dropout(dropout_p, dropout_seed, is_forward=True)
        # Variable-length mode (src/forward/caller.py:25-59): every sequence attends within its own unpadded prefix and
        # padded rows of O are zero. Instead of the reference's pack / unpack copies each sequence is a strided view
        # handed to the kernel (TMA takes the strides); O is written in place.
        assert q.size(1) == k.size(1), "Attention mask is not supported with seqlen_q != seqlen_k"
        assert attention_mask.shape == (batch, seqlen_q), f"{attention_mask.shape = } <> {(batch, seqlen_q) = }"
        scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
        o = torch.zeros_like(q)
        seqlen_q_rounded = math.ceil(seqlen_q / 128) * 128
        lse = torch.zeros((batch, nheads_q, seqlen_q_rounded), device=q.device, dtype=torch.float32)
        if varlen_in_kernel(q, k, v) and not with_dropout:
            # One launch over the whole padded batch: the lengths stay on the device (no .item() / .tolist() sync) and a
            # setup kernel derives per-sequence tensor maps from them (include/fa2_sm100.h, fa2_sm100_fwd_varlen).
            assert nheads_q % nheads_kv == 0, f"{nheads_q = } is not divisible by {nheads_kv =}"
            assert k.shape == expected_kv_shape and v.shape == expected_kv_shape
            assert q.dtype == k.dtype == v.dtype, "All tensors must have the same type"
            assert q.is_cuda and k.is_cuda and v.is_cuda and attention_mask.is_cuda
            lib = _lib.load()
            with torch.cuda.device(q.device):
                seqlens = device_sequence_lengths(attention_mask)
                tmap_ws = torch.empty(lib.fa2_sm100_varlen_workspace_bytes(batch), device=q.device, dtype=torch.uint8)
                code = lib.fa2_sm100_fwd_varlen(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                    batch, nheads_q, nheads_kv, seqlen_q, head_dim,
                    _lib.strides3(q), _lib.strides3(k), _lib.strides3(v), _lib.strides3(o),
                    seqlen_q_rounded, scale, int(bool(causal)), encode_dtype(q),
                    seqlens.data_ptr(), tmap_ws.data_ptr(), tmap_ws.numel(), _stream_ptr(q.device))
            _lib.check(code, "fa2_sm100_fwd_varlen")
            return o, lse, scale, 0 if dropout_seed is None else dropout_seed
        # One launch per sequence over views (odd head dims; dropout). With dropout the reference numbers the counters of
        # a sequence of length L that starts at row c of the PACKED batch as L * (c + L * head) + row * L + col
        # (src/forward/kernel.py:105-112, 147-149: off_batch is 0 in its VARLEN mode): a [1, L, H, D] call whose counters
        # start at L * c draws exactly that mask. The lengths are read on the host here, like src/utils.py:14 does.
        start = 0
        for b, n in enumerate(sequence_lengths(attention_mask)):
            if n == 0:
                continue
            ob, lb, _, _ = _flash_attn_forward(q[b:b + 1, :n], k[b:b + 1, :n], v[b:b + 1, :n], None, None,
                                               dropout_p if with_dropout else 0.0, causal, scale, dropout_seed,
                                               out=o[b:b + 1, :n], dropout_offset=(n * start) & 0xffffffff)
            lse[b, :, :lb.shape[-1]] = lb[0]
            start += n
        return o, lse, scale, 0 if dropout_seed is None else dropout_seed
    if not runs_natively(head_dim):
        d = padded_head_dim(head_dim)
        scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
        o, lse, _, seed = _flash_attn_forward(pad_last(q, d), pad_last(k, d), pad_last(v, d), None, bias, dropout_p,
                                              causal, scale, dropout_seed, dropout_offset=dropout_offset)
        o = o[..., :head_dim]
        if out is not None:
            out.copy_(o)
            o = out
        return o, lse, scale, seed
    # same checks, same exception type as src/forward/caller.py:36-41
    assert nheads_q % nheads_kv == 0, f"{nheads_q = } is not divisible by {nheads_kv =}"
    assert k.shape == expected_kv_shape, f"{k.shape = } <> {expected_kv_shape = }"
    assert v.shape == expected_kv_shape, f"{v.shape = } <> {expected_kv_shape = }"
    assert q.dtype == k.dtype == v.dtype, "All tensors must have the same type"
    assert q.dtype in [torch.float16, torch.bfloat16], "Only support fp16 and bf16"
    assert q.is_cuda and k.is_cuda and v.is_cuda
    softmax_scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
    # bias and dropout (src/forward/caller.py:66-67)
    bias = check_bias(bias, q, batch, nheads_q, seqlen_q, seqlen_k)
    dropout_seed = handle_dropout(dropout_p, dropout_seed, is_forward=True)

    lib = _lib.load()
    with torch.cuda.device(q.device):
        # every element of O and of the padded LSE is written by the kernel: no zero fill needed
        o = torch.empty_like(q) if out is None else out
        assert o.shape == q.shape and o.dtype == q.dtype and o.stride(-1) == 1
        seqlen_q_rounded = math.ceil(seqlen_q / 128) * 128
        lse = torch.empty((batch, nheads_q, seqlen_q_rounded), device=q.device, dtype=torch.float32)
        args = (q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
                batch, nheads_q, nheads_kv, seqlen_q, seqlen_k, head_dim,
                _lib.strides3(q), _lib.strides3(k), _lib.strides3(v), _lib.strides3(o),
                seqlen_q_rounded, softmax_scale, int(bool(causal)), encode_dtype(q))
        if bias is None and not dropout_p:
            code = lib.fa2_sm100_fwd(*args, _stream_ptr(q.device))
        else:
            code = lib.fa2_sm100_fwd_ext(*args, *_lib.bias_args(bias), float(dropout_p), int(dropout_seed),
                                         int(dropout_offset) & 0xffffffff, _stream_ptr(q.device))
    _lib.check(code, "fa2_sm100_fwd")
    return o, lse, softmax_scale, dropout_seed
