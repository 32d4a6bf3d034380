"""Host caller of the backward kernels - the restatement of ``_flash_attn_backward``
(reference src/backward/caller.py:14-178): delta preprocess, fused dK/dV/dQ kernel, dQ conversion."""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .forward import (FUSED_HEAD_DIMS, _stream_ptr, check_bias, device_sequence_lengths, encode_dtype, handle_dropout,
                      kernel_addressable, pad_last, padded_head_dim, runs_natively, sequence_lengths, varlen_in_kernel)

# Process-wide default for how dQ contributions are summed.
#   True     : deterministic - dQ comes from its own kernel, one CTA per query block walking the key blocks in a
#              fixed order, no atomics (what the reference's kernel does, src/backward/kernel.py:154-182). The dK/dV
#              kernel spills its 16-bit dS tiles to a workspace (32 KB per visible 128 x 128 block pair) and the dQ
#              kernel only sums dS K over them: 5 matrix products per block pair (mode 5). If that workspace does
#              not fit (FA2_SPILL_MAX_GIB, default 72, or an allocation failure) the dQ kernel recomputes S, P, dP and
#              dS like the reference (7 products, 256 bytes of workspace; mode 1). Head dim 64 always recomputes: a
#              dS tile is 32 KB whatever the head dim while the products it saves are half as large, and measured
#              (B=2, H=16, S=16384, D=64 non-causal, 20 launches back to back) recomputing takes 8.34 ms against 9.07.
#   1 / 5    : force the recomputing / the spilling deterministic path.
#   False / 0: atomic fast mode - the dK/dV kernel also forms dQ partial products and adds them into an fp32
#              accumulator in arrival order (5 products per block pair).
#   2        : the fused kernel with the adds serialised per query tile by a ticket (bit-repeatable as well).
_DETERMINISTIC = True


def set_deterministic(flag) -> None:
    global _DETERMINISTIC
    _DETERMINISTIC = flag if flag in (0, 1, 2, 5) and not isinstance(flag, bool) else bool(flag)


def _spill_limit_bytes() -> int:
    return int(float(os.environ.get("FA2_SPILL_MAX_GIB", "72")) * (1 << 30))


def is_deterministic() -> bool:
    return _DETERMINISTIC


def _backward_varlen(dO, q, k, v, attention_mask, o, lse, causal, softmax_scale, dq, dk, dv) -> bool:
    """Variable-length backward in one launch per kernel over the padded batch (fa2_sm100_bwd_varlen): the deterministic
    spilled-dS path (or dK / dV alone when dq is None). Returns False - the caller then falls back to one call per
    sequence - when the dS workspace of the padded shape does not fit. dq / dk / dv arrive zero-filled."""
    batch, seqlen, nheads_q, head_dim = q.shape
    if head_dim > 128:
        return False  # the backward kernels for head dims 129..256 take one sequence per launch
    nheads_kv = k.shape[2]
    seqlen_rounded = math.ceil(seqlen / 128) * 128
    scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
    assert nheads_q % nheads_kv == 0, f"{nheads_q = } is not divisible by {nheads_kv =}"
    assert lse.shape == (batch, nheads_q, seqlen_rounded) and lse.dtype == torch.float32 and lse.is_contiguous()
    assert dO.shape == q.shape and o.shape == q.shape and k.shape == v.shape
    assert q.is_cuda and k.is_cuda and v.is_cuda and o.is_cuda and dO.is_cuda and lse.is_cuda and attention_mask.is_cuda
    assert q.dtype == k.dtype == v.dtype == o.dtype == dO.dtype, "All tensors must have the same type"
    assert q.stride(-1) == k.stride(-1) == v.stride(-1) == o.stride(-1) == 1
    lib = _lib.load()
    dt = encode_dtype(q)
    mode = _lib.DQ_DET_SPILL if dq is not None else _lib.DQ_NONE
    ws_bytes = lib.fa2_sm100_bwd_workspace_bytes(batch, nheads_q, seqlen, seqlen, head_dim, int(bool(causal)), mode)
    if ws_bytes > _spill_limit_bytes():
        return False
    with torch.cuda.device(q.device):
        try:
            workspace = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
        except torch.OutOfMemoryError:
            return False
        stream = _stream_ptr(q.device)
        seqlens = device_sequence_lengths(attention_mask)
        tmap_ws = torch.empty(lib.fa2_sm100_varlen_workspace_bytes(batch), device=q.device, dtype=torch.uint8)
        delta = torch.empty_like(lse)
        # dK / dV are always produced by the kernel pair; a caller that only wants dQ gets scratch ones
        dk_ = dk if dk is not None else torch.zeros_like(k)
        dv_ = dv if dv is not None else torch.zeros_like(v)
        code = lib.fa2_sm100_bwd_delta_varlen(
            o.data_ptr(), dO.data_ptr(), delta.data_ptr(), batch, nheads_q, seqlen, head_dim,
            _lib.strides3(o), _lib.strides3(dO), seqlen_rounded, dt, seqlens.data_ptr(), stream)
        _lib.check(code, "fa2_sm100_bwd_delta_varlen")
        code = lib.fa2_sm100_bwd_varlen(
            dO.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(),
            dq.data_ptr() if dq is not None else None, dk_.data_ptr(), dv_.data_ptr(), workspace.data_ptr(), ws_bytes,
            batch, nheads_q, nheads_kv, seqlen, head_dim,
            _lib.strides3(dO), _lib.strides3(q), _lib.strides3(k), _lib.strides3(v),
            _lib.strides3(dq if dq is not None else q), _lib.strides3(dk_), _lib.strides3(dv_),
            seqlen_rounded, scale, int(bool(causal)), mode, dt, encode_dtype(dq) if dq is not None else dt,
            seqlens.data_ptr(), tmap_ws.data_ptr(), tmap_ws.numel(), stream)
        _lib.check(code, "fa2_sm100_bwd_varlen")
    return True


def _flash_attn_backward(
    dO: Tensor,  # [batch_size, seqlen_q, nheads_q, head_dim]
    q: Tensor,  # [batch_size, seqlen_q, nheads_q, head_dim]
    k: Tensor,  # [batch_size, seqlen_k, nheads_kv, head_dim]
    v: Tensor,  # [batch_size, seqlen_k, nheads_kv, head_dim]
    bias: Optional[Tensor],
    attention_mask: Optional[Tensor],
    o: Tensor,  # [batch_size, seqlen_q, nheads_q, head_dim]
    lse: Tensor,  # [batch_size, nheads_q, max_seqlen_q_rounded]
    dropout_p: float,
    causal: bool,
    softmax_scale: Optional[float],
    dropout_seed: Optional[int],
    deterministic: Optional[bool] = None,
    dq_dtype: Optional[torch.dtype] = None,
    need_dq: bool = True,
    need_dkdv: bool = True,
) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    """Returns ``(dq, dk, dv)`` like src/backward/caller.py:178: ``dq`` is fp32 ``[B, Sq, Hq, D]``
    (src/backward/caller.py:86), ``dk``/``dv`` are ``[B, Sk, Hkv, D]`` in the input dtype with the GQA
    head-group sum (src/backward/caller.py:163-165) already applied inside the kernel.

    ``dq_dtype`` (extension): write dq in that dtype instead of fp32. The autograd wrapper passes the input
    dtype - the rounding autograd would apply to the reference's fp32 dq anyway - and saves one pass over dq.
    ``need_dq`` / ``need_dkdv`` (extension): skip the kernel whose outputs autograd does not ask for
    (``ctx.needs_input_grad``); the skipped gradients are returned as ``None``."""
    handle_dropout(dropout_p, dropout_seed, is_forward=False)  # raises NotImplementedError under dropout, like src/utils.py:87-88
    dO = dO.contiguous() if dO.stride(-1) != 1 else dO
    if q.is_cuda and runs_natively(q.shape[-1]):
        # what TMA cannot address (expanded tensors, e.g. the dO of out.sum(); unaligned views) is materialised
        dO, q, k, v, o = (kernel_addressable(t) for t in (dO, q, k, v, o))
    if attention_mask is not None:
        assert bias is None, "Attention mask is not supported along with attention bias. Just use bias instead."
        # Variable-length mode (src/backward/caller.py:29-79, 168-176): per-sequence views instead of pack / unpack;
        # gradients of padded positions are zero.
        assert q.size(1) == k.size(1), "Attention mask is not supported with seqlen_q != seqlen_k"
        dq = torch.zeros_like(q, dtype=torch.float32 if dq_dtype is None else dq_dtype) if need_dq else None
        dk = torch.zeros_like(k) if need_dkdv else None
        dv = torch.zeros_like(v) if need_dkdv else None
        if not need_dq and not need_dkdv:
            return None, None, None
        if varlen_in_kernel(q, k, v):
            done = _backward_varlen(dO, q, k, v, attention_mask, o, lse, causal, softmax_scale, dq, dk, dv)
            if done:
                return dq, dk, dv
        for b, n in enumerate(sequence_lengths(attention_mask)):
            if n == 0:
                continue
            n_pad = math.ceil(n / 128) * 128
            part = _flash_attn_backward(dO[b:b + 1, :n], q[b:b + 1, :n], k[b:b + 1, :n], v[b:b + 1, :n], None, None,
                                        o[b:b + 1, :n], lse[b:b + 1, :, :n_pad].contiguous(), dropout_p, causal,
                                        softmax_scale, dropout_seed, deterministic, dq_dtype, need_dq, need_dkdv)
            for dst, src in zip((dq, dk, dv), part):
                if dst is not None:
                    dst[b, :n] = src[0]
        return dq, dk, dv
    det_req = _DETERMINISTIC if deterministic is None else deterministic
    fused_req = bias is None and (det_req is False or (not isinstance(det_req, bool) and det_req in (0, 2)))
    # zero-padded copies only for head dims TMA cannot address (not a multiple of 8) and for the fused kernel, whose dQ
    # accumulator layout exists for 64 and 128 only
    if fused_req and q.shape[-1] > FUSED_HEAD_DIMS[-1]:
        fused_req, deterministic = False, True  # the fused kernel stops at head dim 128: the deterministic kernels serve it
    if not runs_natively(q.shape[-1]) or (fused_req and q.shape[-1] not in FUSED_HEAD_DIMS):
        head_dim = q.shape[-1]
        d = padded_head_dim(head_dim, FUSED_HEAD_DIMS if fused_req else None)
        scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
        part = _flash_attn_backward(pad_last(dO, d), pad_last(q, d), pad_last(k, d), pad_last(v, d), bias, None,
                                    pad_last(o, d), lse, dropout_p, causal, scale, dropout_seed, deterministic,
                                    dq_dtype, need_dq, need_dkdv)
        return tuple(None if t is None else t[..., :head_dim] for t in part)
    batch_size, seqlen_q, nheads_q, head_dim = q.shape
    _, seqlen_k, nheads_kv, _ = k.shape
    max_seqlen_q_rounded = math.ceil(seqlen_q / 128) * 128
    softmax_scale = 1.0 / math.sqrt(head_dim) if softmax_scale is None else float(softmax_scale)
    # same checks as src/backward/caller.py:51-53
    assert nheads_q % nheads_kv == 0, f"{nheads_q = } is not divisible by {nheads_kv =}"
    assert lse.shape == (batch_size, nheads_q, max_seqlen_q_rounded)
    assert q.stride(-1) == k.stride(-1) == v.stride(-1) == o.stride(-1) == 1
    assert dO.shape == q.shape and o.shape == q.shape
    assert q.is_cuda and k.is_cuda and v.is_cuda and o.is_cuda and dO.is_cuda and lse.is_cuda
    assert q.dtype == k.dtype == v.dtype == o.dtype == dO.dtype, "All tensors must have the same type"
    assert lse.dtype == torch.float32 and lse.is_contiguous()
    bias = check_bias(bias, q, batch_size, nheads_q, seqlen_q, seqlen_k)
    det = _DETERMINISTIC if deterministic is None else deterministic
    if bias is not None and det in (False, 0, 2):
        det = True  # the fused (atomic / ticket-ordered) kernel has no bias variant: the deterministic kernels serve it
    auto = isinstance(det, bool) and det  # True: spill when the workspace fits, else recompute
    if isinstance(det, bool):
        # with a bias the spilled-dS pair is preferred at every head dim: the recomputing dQ kernel has to fetch the bias a
        # second time, row by row (measured at D = 64, config-3 shape: 5.6 ms against 1.6)
        spill = q.shape[-1] > 64 or bias is not None
        mode = (_lib.DQ_DET_SPILL if spill else _lib.DQ_DETERMINISTIC) if det else _lib.DQ_ATOMIC
    else:
        mode = int(det) if det in (0, 1, 2, 5) else int(bool(det))  # one of _lib.DQ_*
    if head_dim > 128 and mode in (_lib.DQ_ATOMIC, _lib.DQ_FUSED_ORDERED, _lib.DQ_DETERMINISTIC):
        # head dims 129..256 have one backward: deterministic, dQ in its own kernel, P / dS tiles spilled for dK / dV
        mode, auto = _lib.DQ_DET_SPILL, False
    if not need_dq and not need_dkdv:
        return None, None, None
    if not need_dq:
        mode = _lib.DQ_NONE
    elif not need_dkdv:
        mode = _lib.DQ_ONLY

    lib = _lib.load()
    dt = encode_dtype(q)
    with torch.cuda.device(q.device):
        stream = _stream_ptr(q.device)
        delta = torch.empty_like(lse)
        dq = torch.empty_like(q, dtype=torch.float32 if dq_dtype is None else dq_dtype) if need_dq else None
        dk = torch.empty_like(k) if need_dkdv else None
        dv = torch.empty_like(v) if need_dkdv else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        s3 = lambda t, like: _lib.strides3(t if t is not None else like)
        fused = mode in (_lib.DQ_ATOMIC, _lib.DQ_FUSED_ORDERED)
        # the _gqa size leaves room for the fp32 partials of split head groups (MQA / large GQA groups)
        ws_args = (batch_size, nheads_q, nheads_kv, seqlen_q, seqlen_k, head_dim, int(bool(causal)))
        ws_bytes = lib.fa2_sm100_bwd_workspace_bytes_gqa(*ws_args, mode)
        workspace = None
        if auto and mode == _lib.DQ_DET_SPILL:
            if ws_bytes <= _spill_limit_bytes():
                try:
                    workspace = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
                except torch.OutOfMemoryError:
                    workspace = None
            if workspace is None:  # no room for the dS tiles: the recomputing dQ kernel needs 256 bytes
                mode = _lib.DQ_DETERMINISTIC
                ws_bytes = lib.fa2_sm100_bwd_workspace_bytes_gqa(*ws_args, mode)
        if workspace is None and ws_bytes:
            workspace = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
        ws_ptr = workspace.data_ptr() if ws_bytes else None

        # delta = rowsum(O * dO); the same launch clears the dQ accumulator when there is one
        code = lib.fa2_sm100_bwd_delta(
            o.data_ptr(), dO.data_ptr(), delta.data_ptr(), batch_size, nheads_q, seqlen_q, head_dim,
            _lib.strides3(o), _lib.strides3(dO), max_seqlen_q_rounded, dt,
            ws_ptr if fused else None, ws_bytes if fused else 0, stream)
        _lib.check(code, "fa2_sm100_bwd_delta")
        code = lib.fa2_sm100_bwd_ext(
            dO.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(),
            ptr(dq), ptr(dk), ptr(dv), ws_ptr, ws_bytes,
            batch_size, nheads_q, nheads_kv, seqlen_q, seqlen_k, head_dim,
            _lib.strides3(dO), _lib.strides3(q), _lib.strides3(k), _lib.strides3(v),
            s3(dq, q), s3(dk, k), s3(dv, v),
            max_seqlen_q_rounded, softmax_scale, int(bool(causal)), mode, dt,
            encode_dtype(dq) if dq is not None else dt, *_lib.bias_args(bias), stream)
        _lib.check(code, "fa2_sm100_bwd")
    return dq, dk, dv
