"""Host caller of the forward kernel - the restatement of ``_flash_attn_forward``
(reference src/forward/caller.py:12-122): validate, allocate O and LSE, launch."""
from __future__ import annotations

import math
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


def _reject_out_of_scope(attention_mask, bias, dropout_p):
    # varlen masks, additive bias and dropout are outside the hot path of this build.
    if attention_mask is not None:
        raise NotImplementedError("attention_mask (varlen packing) is not part of the B200 hot path yet")
    if bias is not None:
        raise NotImplementedError("attention_bias is not part of the B200 hot path yet")
    if dropout_p and dropout_p > 0:
        raise NotImplementedError("dropout is not part of the B200 hot path yet")


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
) -> Tuple[Tensor, Tensor, float, int]:
    """Returns ``(o, lse, softmax_scale, dropout_seed)`` like src/forward/caller.py:122.

    ``lse`` is fp32 ``[B, Hq, ceil(Sq/128)*128]`` in the log2 domain (src/forward/caller.py:73-74,
    src/forward/kernel.py:119); entries past ``Sq`` are 0.
    """
    _reject_out_of_scope(attention_mask, bias, dropout_p)
    batch, seqlen_q, nheads_q, head_dim = q.shape
    _, seqlen_k, nheads_kv, _ = k.shape
    expected_kv_shape = (batch, seqlen_k, nheads_kv, head_dim)
    # same checks, same exception type as src/forward/caller.py:36-41
    assert nheads_q % nheads_kv == 0, f"{nheads_q = } is not divisible by {nheads_kv =}"
    assert k.shape == expected_kv_shape, f"{k.shape = } <> {expected_kv_shape = }"
    assert v.shape == expected_kv_shape, f"{v.shape = } <> {expected_kv_shape = }"
    assert q.dtype == k.dtype == v.dtype, "All tensors must have the same type"
    assert q.dtype in [torch.float16, torch.bfloat16], "Only support fp16 and bf16"
    assert q.is_cuda and k.is_cuda and v.is_cuda
    softmax_scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
    dropout_seed = 0 if dropout_seed is None else dropout_seed

    lib = _lib.load()
    with torch.cuda.device(q.device):
        # every element of O and of the padded LSE is written by the kernel: no zero fill needed
        o = torch.empty_like(q)
        seqlen_q_rounded = math.ceil(seqlen_q / 128) * 128
        lse = torch.empty((batch, nheads_q, seqlen_q_rounded), device=q.device, dtype=torch.float32)
        code = lib.fa2_sm100_fwd(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(),
            batch, nheads_q, nheads_kv, seqlen_q, seqlen_k, head_dim,
            _lib.strides3(q), _lib.strides3(k), _lib.strides3(v), _lib.strides3(o),
            seqlen_q_rounded, softmax_scale, int(bool(causal)), encode_dtype(q), _stream_ptr(q.device))
    _lib.check(code, "fa2_sm100_fwd")
    return o, lse, softmax_scale, dropout_seed
