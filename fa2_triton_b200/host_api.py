"""Host-buffer entry point: attention forward + backward for tensors that live in (pinned) host memory.

This is what a caller that keeps activations off-device binds to, and what ``bench.py`` times as ``e2e``: every
call moves q, k, v, dO host->device, runs the four kernels, and moves O, dQ, dK, dV device->host. The batch is cut
into chunks that flow through three streams (upload / compute / download), so the PCIe transfers of neighbouring
chunks overlap each other (the link is full duplex) and the kernels.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .backward import _flash_attn_backward
from .forward import _flash_attn_forward


class HostAttention:
    """Reusable plan for one problem shape: device staging buffers, pinned result buffers and streams."""

    def __init__(self, q_shape, kv_shape, dtype=torch.bfloat16, device="cuda:0", chunks: Optional[int] = None):
        assert torch.cuda.is_available(), "HostAttention needs a CUDA device; there is no CPU fallback"
        self.device = torch.device(device)
        B = q_shape[0]
        self.chunks = min(B, chunks) if chunks else B
        self.bounds = [(i * B // self.chunks, (i + 1) * B // self.chunks) for i in range(self.chunks)]
        dev = dict(device=self.device, dtype=dtype)
        self.d: Dict[str, Tensor] = {
            "q": torch.empty(q_shape, **dev), "k": torch.empty(kv_shape, **dev), "v": torch.empty(kv_shape, **dev),
            "do": torch.empty(q_shape, **dev),
        }
        pin = dict(dtype=dtype, pin_memory=True)
        self.h: Dict[str, Tensor] = {
            "o": torch.empty(q_shape, **pin), "dq": torch.empty(q_shape, **pin),
            "dk": torch.empty(kv_shape, **pin), "dv": torch.empty(kv_shape, **pin),
        }
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.d.values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.h.values())

    def __call__(self, q: Tensor, k: Tensor, v: Tensor, do: Tensor, causal: bool = False,
                 softmax_scale: Optional[float] = None, deterministic: Optional[bool] = None
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """q, k, v, do: host tensors (pinned for asynchronous copies). Returns pinned host (o, dq, dk, dv); they
        are valid when the call returns (it synchronises on the download stream)."""
        src = {"q": q, "k": k, "v": v, "do": do}
        caller = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(caller)
        keep = []
        for (b0, b1) in self.bounds:
            with torch.cuda.stream(self.s_in):
                for name in ("q", "k", "v", "do"):
                    self.d[name][b0:b1].copy_(src[name][b0:b1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ready)
                qc, kc, vc, doc = (self.d[n][b0:b1] for n in ("q", "k", "v", "do"))
                o, lse, scale, _ = _flash_attn_forward(qc, kc, vc, None, None, 0.0, causal, softmax_scale, None)
                dq, dk, dv = _flash_attn_backward(doc, qc, kc, vc, None, None, o, lse, 0.0, causal, scale, None,
                                                  deterministic=deterministic, dq_dtype=qc.dtype)
                done = torch.cuda.Event()
                done.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                for name, t in (("o", o), ("dq", dq), ("dk", dk), ("dv", dv)):
                    self.h[name][b0:b1].copy_(t, non_blocking=True)
                    t.record_stream(self.s_out)
            keep.append((o, dq, dk, dv, lse))
        caller.wait_stream(self.s_out)
        self.s_out.synchronize()
        return self.h["o"], self.h["dq"], self.h["dk"], self.h["dv"]


_plans: Dict[tuple, HostAttention] = {}


def flash_attn_fwd_bwd_host(q: Tensor, k: Tensor, v: Tensor, do: Tensor, causal: bool = False,
                            softmax_scale: Optional[float] = None, deterministic: Optional[bool] = None,
                            device="cuda:0", chunks: Optional[int] = None):
    """One-shot convenience wrapper around :class:`HostAttention` (plans are cached per shape)."""
    key = (tuple(q.shape), tuple(k.shape), q.dtype, str(device), chunks)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = HostAttention(q.shape, k.shape, q.dtype, device, chunks)
    return plan(q, k, v, do, causal, softmax_scale, deterministic)
