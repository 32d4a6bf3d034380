"""The autograd entry point - same names, argument order and return convention as the reference's
``src/wrapper.py`` (FlashAttnFunc :10-86, flash_attn_func :89-100), so it drops in for it."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .backward import _flash_attn_backward
from .forward import _flash_attn_forward


class FlashAttnFunc(torch.autograd.Function):

    @staticmethod
    def forward(
        ctx,
        q: Tensor,
        k: Tensor,
        v: Tensor,
        attention_mask: Optional[Tensor] = None,
        attention_bias: Optional[Tensor] = None,
        dropout_p: float = 0.0,
        causal: bool = False,
        softmax_scale: Optional[float] = None,
        dropout_seed: Optional[int] = None,
    ):
        """q: [B, Sq, Hq, D]; k, v: [B, Sk, Hkv, D]; returns O [B, Sq, Hq, D] (src/wrapper.py:13-59)."""
        # last dimension must be contiguous (src/wrapper.py:41-43)
        q = q if q.stride(-1) == 1 else q.contiguous()
        k = k if k.stride(-1) == 1 else k.contiguous()
        v = v if v.stride(-1) == 1 else v.contiguous()
        attention_bias = None if (attention_bias is None) else attention_bias.contiguous()
        o, lse, ctx.softmax_scale, ctx.dropout_seed = _flash_attn_forward(
            q=q,
            k=k,
            v=v,
            attention_mask=attention_mask,
            bias=attention_bias,
            dropout_p=dropout_p,
            causal=causal,
            softmax_scale=softmax_scale,
            dropout_seed=dropout_seed,
        )
        ctx.save_for_backward(q, k, v, attention_bias, attention_mask, o, lse)
        ctx.causal = causal
        ctx.dropout_p = dropout_p
        return o

    @staticmethod
    def backward(ctx, do):
        """do: [B, Sq, Hq, D]; returns (dq, dk, dv, None x 6) (src/wrapper.py:62-86)."""
        q, k, v, bias, attention_mask, o, lse = ctx.saved_tensors
        dq, dk, dv = _flash_attn_backward(
            dO=do,
            q=q,
            k=k,
            v=v,
            bias=bias,
            attention_mask=attention_mask,
            o=o,
            lse=lse,
            dropout_p=ctx.dropout_p,
            causal=ctx.causal,
            softmax_scale=ctx.softmax_scale,
            dropout_seed=ctx.dropout_seed,
            dq_dtype=q.dtype,
            need_dq=ctx.needs_input_grad[0],
            need_dkdv=ctx.needs_input_grad[1] or ctx.needs_input_grad[2],
        )
        return dq, dk, dv, None, None, None, None, None, None


def flash_attn_func(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    attention_mask: Optional[Tensor] = None,
    attention_bias: Optional[Tensor] = None,
    dropout_p: float = 0.0,
    causal: bool = False,
    softmax_scale: Optional[float] = None,
    dropout_seed: Optional[int] = None,
) -> Tensor:
    return FlashAttnFunc.apply(q, k, v, attention_mask, attention_bias, dropout_p, causal, softmax_scale, dropout_seed)
