"""Eager-PyTorch port of the reference's CPU path. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``flash_attn_reference`` restates src/reference_implementation.py:38-123 for the arguments the hot path uses
(q, k, v, causal, upcast, reorder_ops): einsum scores -> -inf masking -> softmax -> einsum with V, gradients by
PyTorch autograd. It is the thing ``bench.py`` times as the CPU baseline (kind "port") and the "low precision
PyTorch" arm of the reference's error criterion (tests/utils.py:68-142). It is validated against the numpy
oracle and the golden vectors in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import torch


def flash_attn_reference(q, k, v, causal=False, upcast=True, reorder_ops=False, softmax_scale=None, padding_mask=None,
                         attn_bias=None, dropout_p=0.0, dropout_mask=None):
    """q [B,Sq,Hq,D], k/v [B,Sk,Hkv,D] -> output [B,Sq,Hq,D] in q's dtype.

    padding_mask: optional bool [B, S] used as both query_padding_mask and key_padding_mask, the way the reference's
    tests call it (tests/core.py:31-49): padded keys are hidden (:91-92), the causal diagonal follows the per-sequence
    lengths (:16-29), padded query rows produce zeros (:111-112, :121-122).

    attn_bias: broadcastable to [B, Hq, Sq, Sk], added to the scaled, masked scores (:101-102). dropout_mask: bool
    [B, Hq, Sq, Sk] keep-mask applied to the softmax output, V scaled by 1 / (1 - dropout_p) (:113-120).

    upcast: compute in fp32 (src/reference_implementation.py:77-78). reorder_ops: scale K instead of Q
    (:83-86), used by the reference tests to estimate the error of a same-precision implementation.
    """
    out_dtype = q.dtype
    if upcast:
        q, k, v = q.float(), k.float(), v.float()
    sq, sk = q.shape[1], k.shape[1]
    group = q.shape[2] // k.shape[2]
    if group > 1:  # :80-81
        k = k.repeat_interleave(group, dim=2)
        v = v.repeat_interleave(group, dim=2)
    scale = 1.0 / math.sqrt(q.shape[-1]) if softmax_scale is None else softmax_scale
    if reorder_ops:
        scores = torch.einsum("bthd,bshd->bhts", q, k * scale)
    else:
        scores = torch.einsum("bthd,bshd->bhts", q * scale, k)
    hidden = None
    if padding_mask is not None:
        scores = scores.masked_fill((~padding_mask).unsqueeze(1).unsqueeze(1), float("-inf"))
    if causal:  # :8-29 with window (-1, 0): bottom-right aligned
        row = torch.arange(sq, device=q.device).unsqueeze(-1)
        col = torch.arange(sk, device=q.device)
        if padding_mask is None:
            hidden = col > row + (sk - sq)
        else:  # per-sequence lengths; query and key lengths coincide, so the diagonal is the main one
            hidden = (col > row).expand(q.shape[0], 1, sq, sk)
        scores = scores.masked_fill(hidden, float("-inf"))
    if attn_bias is not None:
        scores = scores + attn_bias.to(scores.dtype)
    attn = torch.softmax(scores, dim=-1).to(v.dtype)  # :105
    if hidden is not None:  # rows without any visible key: 0 instead of NaN (:107-108)
        attn = attn.masked_fill(torch.all(hidden, dim=-1, keepdim=True), 0.0)
    if padding_mask is not None:
        attn = attn.masked_fill((~padding_mask).unsqueeze(1).unsqueeze(-1), 0.0)
    if dropout_mask is not None:
        attn = attn.masked_fill(~dropout_mask, 0.0)
    out = torch.einsum("bhts,bshd->bthd", attn, v * (1.0 / (1.0 - dropout_p)))  # :120
    if padding_mask is not None:
        out = out.masked_fill((~padding_mask).unsqueeze(-1).unsqueeze(-1), 0.0)
    return out.to(out_dtype)


def fwd_bwd(q, k, v, do, causal=False, upcast=True, reorder_ops=False, padding_mask=None, softmax_scale=None,
            attn_bias=None):
    """Forward + autograd backward: (out, dq, dk, dv). Inputs are left untouched."""
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = flash_attn_reference(q, k, v, causal=causal, upcast=upcast, reorder_ops=reorder_ops, padding_mask=padding_mask,
                               softmax_scale=softmax_scale, attn_bias=attn_bias)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    return out.detach(), dq, dk, dv


def lse_log2(q, k, causal=False, softmax_scale=None, attn_bias=None):
    """fp32 log2-domain logsumexp [B, Hq, Sq] - the quantity the forward kernel saves
    (src/forward/kernel.py:119; tests/test_logsumexp.py:53-74 states the same definition)."""
    q, k = q.float(), k.float()
    sq, sk = q.shape[1], k.shape[1]
    group = q.shape[2] // k.shape[2]
    if group > 1:
        k = k.repeat_interleave(group, dim=2)
    scale = 1.0 / math.sqrt(q.shape[-1]) if softmax_scale is None else softmax_scale
    scores = torch.einsum("bthd,bshd->bhts", q, k) * scale
    if causal:
        row = torch.arange(sq, device=q.device).unsqueeze(-1)
        col = torch.arange(sk, device=q.device)
        scores = scores.masked_fill(col > row + (sk - sq), float("-inf"))
    if attn_bias is not None:
        scores = scores + attn_bias.float()
    return torch.logsumexp(scores, dim=-1) * 1.4426950408889634
