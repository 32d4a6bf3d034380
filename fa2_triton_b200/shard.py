"""Batch x head partitioner: shards the independent (batch, head) slices of one attention call across the GPUs of a
box, one process per GPU, with no collective on the data path.

Every (b, h) slice of the forward and of the backward is independent - the reference's grids carry ``B * Hq`` as a
pure parallel axis (src/forward/caller.py:82, src/backward/caller.py:118-121) and no kernel reads across it. So a
G-GPU run is G independent single-GPU runs over disjoint slices; NCCL (``torch.distributed``) is used only by
:func:`gather_for_verification` and for the timing barrier.

Units are kv-head groups: one unit = one batch element x one kv head (with its ``Hq / Hkv`` query heads), so that
the GQA head-group reduction of dK/dV stays inside one GPU. Ranks receive contiguous unit ranges that differ in
size by at most one; ranges are expressed as per-batch head spans so they can be cut out of ``[B, S, H, D]``
tensors as views (``narrow`` along batch and head: TMA handles the resulting strides, nothing is copied).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor


@dataclass(frozen=True)
class Span:
    """Heads ``[kv_head0, kv_head1)`` of batch element ``batch``."""
    batch: int
    kv_head0: int
    kv_head1: int

    @property
    def units(self) -> int:
        return self.kv_head1 - self.kv_head0


def unit_range(n_units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by <= 1) range of units owned by ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def partition(batch: int, nheads_kv: int, world_size: int, rank: int) -> List[Span]:
    """Spans owned by ``rank``; units are numbered ``b * nheads_kv + h_kv``."""
    lo, hi = unit_range(batch * nheads_kv, world_size, rank)
    spans: List[Span] = []
    u = lo
    while u < hi:
        b, h = divmod(u, nheads_kv)
        h_end = min(nheads_kv, h + (hi - u))
        spans.append(Span(b, h, h_end))
        u += h_end - h
    return spans


def merge_spans(spans: Sequence[Span], nheads_kv: int) -> List[Tuple[int, int, int, int]]:
    """Fuse runs of whole-batch spans into ``(batch0, batch1, kv_head0, kv_head1)`` blocks, so that a rank owning
    several complete batch elements launches once over all of them."""
    blocks: List[Tuple[int, int, int, int]] = []
    for s in spans:
        whole = s.kv_head0 == 0 and s.kv_head1 == nheads_kv
        if whole and blocks and blocks[-1][2] == 0 and blocks[-1][3] == nheads_kv and blocks[-1][1] == s.batch:
            b0, _, h0, h1 = blocks[-1]
            blocks[-1] = (b0, s.batch + 1, h0, h1)
        else:
            blocks.append((s.batch, s.batch + 1, s.kv_head0, s.kv_head1))
    return blocks


def shard_views(q: Tensor, k: Tensor, v: Tensor, world_size: int, rank: int, *extra_q_like: Tensor):
    """Views of this rank's slices. q and ``extra_q_like`` (o, do, dq...) are ``[B, Sq, Hq, D]``; k, v are
    ``[B, Sk, Hkv, D]``. Returns a list of ``(block, q_view, k_view, v_view, *extra_views)``."""
    B, _, Hq, _ = q.shape
    Hkv = k.shape[2]
    if Hq % Hkv:
        raise ValueError(f"{Hq=} is not divisible by {Hkv=}")
    group = Hq // Hkv
    out = []
    for (b0, b1, h0, h1) in merge_spans(partition(B, Hkv, world_size, rank), Hkv):
        cut_q = lambda t: t[b0:b1, :, h0 * group:h1 * group]
        cut_kv = lambda t: t[b0:b1, :, h0:h1]
        out.append(((b0, b1, h0, h1), cut_q(q), cut_kv(k), cut_kv(v), *(cut_q(t) for t in extra_q_like)))
    return out


def shard_mask_and_bias(block, nheads_q: int, nheads_kv: int, attention_mask: Optional[Tensor],
                        attention_bias: Optional[Tensor]):
    """This block's part of the per-call extras: the ``[B, S]`` attention mask is cut along the batch, a
    ``[1 | B, 1 | Hq, Sq, Sk]`` bias along whichever of its first two dimensions is not broadcast (views)."""
    b0, b1, h0, h1 = block
    group = nheads_q // nheads_kv
    mask = None if attention_mask is None else attention_mask[b0:b1]
    bias = attention_bias
    if bias is not None:
        if bias.size(0) > 1:
            bias = bias[b0:b1]
        if bias.size(1) > 1:
            bias = bias[:, h0 * group:h1 * group]
    return mask, bias


def sharded_attention(q: Tensor, k: Tensor, v: Tensor, world_size: int, rank: int, causal: bool = False,
                      softmax_scale: Optional[float] = None, attn: Optional[Callable] = None,
                      attention_mask: Optional[Tensor] = None, attention_bias: Optional[Tensor] = None):
    """Run attention on this rank's (batch, head) slices of globally-shaped q, k, v (every rank holds, or can
    view, the tensors it owns; slices of other ranks are never touched). Returns ``[(block, out_view), ...]``.
    ``attn`` defaults to :func:`fa2_triton_b200.flash_attn_func`; tests substitute a CPU function.
    ``attention_mask`` / ``attention_bias`` are sliced with the tensors. Dropout is not offered here: the reference's
    mask is numbered over the whole ``[B, Hq, Sq, Sk]`` score tensor (src/forward/kernel.py:147-149), a shard would
    draw a different one."""
    if attn is None:
        from .wrapper import flash_attn_func as attn
    results = []
    Hq, Hkv = q.shape[2], k.shape[2]
    for block, qs, ks, vs in shard_views(q, k, v, world_size, rank):
        kw = {}
        if attention_mask is not None or attention_bias is not None:
            mask, bias = shard_mask_and_bias(block, Hq, Hkv, attention_mask, attention_bias)
            kw = dict(attention_mask=mask, attention_bias=bias)
        results.append((block, attn(qs, ks, vs, causal=causal, softmax_scale=softmax_scale, **kw)))
    return results


def gather_for_verification(local: Tensor, global_shape: Sequence[int], blocks, group=None, heads_per_unit: int = 1,
                            dtype: Optional[torch.dtype] = None, device=None):
    """All-gather every rank's output blocks into a full ``global_shape`` tensor on every rank. Verification only -
    this is the single place the multi-GPU path uses a collective.

    ``local`` is a list of tensors matching ``blocks`` (this rank's ``(b0, b1, h0, h1)`` blocks, kv-head units;
    ``heads_per_unit`` = Hq / Hkv for q-like tensors, 1 for kv-like). ``dtype`` / ``device`` of the gathered tensor:
    taken from ``local`` when this rank has blocks; a rank without work MUST be told (every rank's broadcast buffers
    have to agree in dtype with the senders'), otherwise it raises."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if local:
        dtype = local[0].dtype if dtype is None else dtype
        device = local[0].device if device is None else device
    if dtype is None:
        raise ValueError("gather_for_verification: a rank without blocks needs the dtype of the gathered tensor")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    full = torch.zeros(*global_shape, dtype=dtype, device=device)
    meta = [None] * world
    dist.all_gather_object(meta, [tuple(b) for b in blocks], group=group)
    for src in range(world):
        for i, (b0, b1, h0, h1) in enumerate(meta[src]):
            view = full[b0:b1, :, h0 * heads_per_unit:h1 * heads_per_unit]
            buf = local[i].to(dtype).contiguous() if src == dist.get_rank(group) else torch.empty_like(view.contiguous())
            dist.broadcast(buf, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
            view.copy_(buf)
    return full
