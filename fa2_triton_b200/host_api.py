"""Host-buffer entry point: attention forward + backward for tensors that live in (pinned) host memory.

This is what a caller that keeps activations off-device binds to, and what ``bench.py`` times as ``e2e``: every
call moves q, k, v, dO host->device, runs the four kernels, and moves O, dQ, dK, dV device->host. The work is cut
into (batch element, kv-head range) chunks - independent units of attention - that flow
through three streams (upload / compute / download), so the PCIe transfers of neighbouring chunks overlap each other
(the link is full duplex) and the kernels. A chunk is a pitched slice of the [B, S, H, D] host tensors: it is moved
with cudaMemcpy2DAsync (fa2_sm100_copy2d_async) straight between the caller's pinned buffers and the device tensors,
never packed. The link bounds the call (config 3: 512 MiB each way), so what matters is the pipeline fill - the upload
of the first chunk and the download of the last - hence ~16 chunks rather than one per batch element.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .backward import _flash_attn_backward
from .forward import _flash_attn_forward


def _copy_slice(dst: Tensor, src: Tensor, to_device: bool, stream: torch.cuda.Stream) -> None:
    """dst / src: equally shaped [S, h, D] slices (rows of h * D contiguous elements, any row pitch)."""
    assert dst.shape == src.shape and dst.stride(-1) == 1 and src.stride(-1) == 1
    assert dst.stride(1) == dst.shape[2] and src.stride(1) == src.shape[2]
    es = dst.element_size()
    code = _lib.load().fa2_sm100_copy2d_async(dst.data_ptr(), dst.stride(0) * es, src.data_ptr(), src.stride(0) * es,
                                              dst.shape[1] * dst.shape[2] * es, dst.shape[0], int(to_device),
                                              stream.cuda_stream)
    _lib.check(code, "fa2_sm100_copy2d_async")


class HostTicket:
    """One submitted forward + backward: ``result()`` blocks until its outputs are in host memory and returns the
    pinned host tensors ``(o, dq, dk, dv)``; they stay valid until ``depth`` further submissions reuse their slot."""

    def __init__(self, outputs, finished: torch.cuda.Event):
        self._outputs, self._finished = outputs, finished

    def result(self) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        self._finished.synchronize()
        return self._outputs


class HostAttention:
    """Reusable plan for one problem shape: device staging buffers, pinned result buffers and streams.

    ``depth`` calls may be in flight at once (``submit`` / ``HostTicket.result``): each has its own device inputs and
    pinned outputs, so the upload of call i + 1 runs while call i computes and downloads - the pipeline fill and drain
    of a call (the upload of its first chunk, the download of its last) overlap with its neighbours' and the link stays
    busy in both directions. ``plan(q, k, v, do)`` is ``submit(...).result()``."""

    def __init__(self, q_shape, kv_shape, dtype=torch.bfloat16, device="cuda:0", chunks: Optional[int] = None,
                 depth: int = 2):
        assert torch.cuda.is_available(), "HostAttention needs a CUDA device; there is no CPU fallback"
        self.device = torch.device(device)
        B, Hq, Hkv = q_shape[0], q_shape[2], kv_shape[2]
        self.group = Hq // Hkv
        want = chunks if chunks else 16
        per_batch = max(1, min(Hkv, want // B))
        while Hkv % per_batch:  # equal kv-head ranges
            per_batch -= 1
        step = Hkv // per_batch
        # (batch element, first kv head, last kv head + 1)
        self.units = [(b, g, g + step) for b in range(B) for g in range(0, Hkv, step)]
        self.chunks = len(self.units)
        self.depth = max(1, int(depth))
        dev = dict(device=self.device, dtype=dtype)
        pin = dict(dtype=dtype, pin_memory=True)
        self.d, self.h = [], []
        for _ in range(self.depth):
            self.d.append({"q": torch.empty(q_shape, **dev), "k": torch.empty(kv_shape, **dev),
                           "v": torch.empty(kv_shape, **dev), "do": torch.empty(q_shape, **dev)})
            self.h.append({"o": torch.empty(q_shape, **pin), "dq": torch.empty(q_shape, **pin),
                           "dk": torch.empty(kv_shape, **pin), "dv": torch.empty(kv_shape, **pin)})
        self._tickets = [None] * self.depth
        self._next = 0
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.d[0].values())
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.h[0].values())

    def submit(self, q: Tensor, k: Tensor, v: Tensor, do: Tensor, causal: bool = False,
               softmax_scale: Optional[float] = None, deterministic: Optional[bool] = None) -> HostTicket:
        """Enqueue one forward + backward for host tensors q, k, v, do (pinned for asynchronous copies; they must stay
        unchanged until the ticket's ``result()`` returns) and return at once."""
        slot = self._next % self.depth
        self._next += 1
        if self._tickets[slot] is not None:  # the slot's previous call must have left the device before it is reused
            self._tickets[slot]._finished.synchronize()
        dbuf, hbuf = self.d[slot], self.h[slot]
        src = {"q": q, "k": k, "v": v, "do": do}
        for name, t in src.items():
            assert t.shape == dbuf[name].shape and t.dtype == dbuf[name].dtype and t.is_contiguous(), name
        caller = torch.cuda.current_stream(self.device)
        self.s_in.wait_stream(caller)
        grp = self.group
        for (b, g0, g1) in self.units:
            heads = {"q": (g0 * grp, g1 * grp), "do": (g0 * grp, g1 * grp), "k": (g0, g1), "v": (g0, g1),
                     "o": (g0 * grp, g1 * grp), "dq": (g0 * grp, g1 * grp), "dk": (g0, g1), "dv": (g0, g1)}
            for name in ("q", "k", "v", "do"):
                h0, h1 = heads[name]
                _copy_slice(dbuf[name][b, :, h0:h1], src[name][b, :, h0:h1], True, self.s_in)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ready)
                qc, kc, vc, doc = (dbuf[n][b:b + 1, :, heads[n][0]:heads[n][1]] for n in ("q", "k", "v", "do"))
                o, lse, scale, _ = _flash_attn_forward(qc, kc, vc, None, None, 0.0, causal, softmax_scale, None)
                dq, dk, dv = _flash_attn_backward(doc, qc, kc, vc, None, None, o, lse, 0.0, causal, scale, None,
                                                  deterministic=deterministic, dq_dtype=qc.dtype)
                done = torch.cuda.Event()
                done.record(self.s_run)
            self.s_out.wait_event(done)
            for name, t in (("o", o), ("dq", dq), ("dk", dk), ("dv", dv)):
                h0, h1 = heads[name]
                _copy_slice(hbuf[name][b, :, h0:h1], t[0], False, self.s_out)
                t.record_stream(self.s_out)
        finished = torch.cuda.Event()
        finished.record(self.s_out)
        ticket = HostTicket((hbuf["o"], hbuf["dq"], hbuf["dk"], hbuf["dv"]), finished)
        self._tickets[slot] = ticket
        return ticket

    def __call__(self, q: Tensor, k: Tensor, v: Tensor, do: Tensor, causal: bool = False,
                 softmax_scale: Optional[float] = None, deterministic: Optional[bool] = None
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """q, k, v, do: host tensors (pinned for asynchronous copies). Returns pinned host (o, dq, dk, dv); they
        are valid when the call returns (it waits for the download stream)."""
        return self.submit(q, k, v, do, causal, softmax_scale, deterministic).result()


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
