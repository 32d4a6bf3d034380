"""CPU tests of the batch x head partitioner (fa2_triton_b200/shard.py): unit ranges, spans, views, and the
world_size-2 path with the gloo backend (no collective on the data path; gather only for verification)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fa2_triton_b200 import shard
from oracle import eager_reference as er


@pytest.mark.parametrize("n,world", [(128, 8), (10, 4), (3, 8), (1024, 7), (1, 1)])
def test_unit_ranges_tile_the_work(n, world):
    ranges = [shard.unit_range(n, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


def test_partition_spans_cover_every_unit_once():
    B, Hkv, world = 5, 6, 4
    seen = set()
    for r in range(world):
        for s in shard.partition(B, Hkv, world, r):
            for h in range(s.kv_head0, s.kv_head1):
                assert (s.batch, h) not in seen
                seen.add((s.batch, h))
    assert len(seen) == B * Hkv


def test_whole_batches_merge_into_one_block():
    # config 5 of BASELINE.json: B=32, H=32 over 8 GPUs -> 4 whole batch elements per rank, one launch each
    for r in range(8):
        blocks = shard.merge_spans(shard.partition(32, 32, 8, r), 32)
        assert blocks == [(4 * r, 4 * r + 4, 0, 32)]


def test_shard_views_are_views_and_respect_gqa():
    B, S, Hq, Hkv, D = 2, 16, 8, 2, 64
    q = torch.randn(B, S, Hq, D)
    k = torch.randn(B, S, Hkv, D)
    v = torch.randn(B, S, Hkv, D)
    pieces = [p for r in range(4) for p in shard.shard_views(q, k, v, 4, r)]
    assert len(pieces) == 4
    for (b0, b1, h0, h1), qs, ks, vs in pieces:
        assert qs.data_ptr() == q[b0:b1, :, h0 * 4:h1 * 4].data_ptr()  # a view, nothing copied
        assert qs.shape[2] == 4 * ks.shape[2]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # every rank can rebuild the global tensors; it only touches its own slices
        B, S, Hq, Hkv, D = 3, 64, 4, 2, 32
        q = torch.randn(B, S, Hq, D)
        k = torch.randn(B, S, Hkv, D)
        v = torch.randn(B, S, Hkv, D)
        attn = lambda a, b, c, causal, softmax_scale: er.flash_attn_reference(a, b, c, causal=causal, softmax_scale=softmax_scale)
        out = shard.sharded_attention(q, k, v, world, rank, causal=True, attn=attn)
        blocks = [b for b, _ in out]
        full = shard.gather_for_verification([o for _, o in out], q.shape, blocks, heads_per_unit=Hq // Hkv, dtype=q.dtype,
                                             device=q.device)
        ref = er.flash_attn_reference(q, k, v, causal=True)
        err = float((full - ref).abs().max())
        # the per-call extras travel with the slices: a per-(batch, head) bias and a right-padded attention mask
        bias = torch.randn(B, Hq, S, S)
        mask = torch.ones(B, S, dtype=torch.bool)
        mask[1, 40:] = False
        for kw_in, kw_ref in ((dict(attention_bias=bias), dict(attn_bias=bias)), (dict(attention_mask=mask), dict(padding_mask=mask))):
            attn2 = lambda a, b, c, causal, softmax_scale, attention_mask=None, attention_bias=None: er.flash_attn_reference(
                a, b, c, causal=causal, softmax_scale=softmax_scale, padding_mask=attention_mask, attn_bias=attention_bias)
            out2 = shard.sharded_attention(q, k, v, world, rank, causal=True, attn=attn2, **kw_in)
            full2 = shard.gather_for_verification([o for _, o in out2], q.shape, [b for b, _ in out2], heads_per_unit=Hq // Hkv,
                                                  dtype=q.dtype, device=q.device)
            err = max(err, float((full2 - er.flash_attn_reference(q, k, v, causal=True, **kw_ref)).abs().max()))
        results[rank] = err
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_match_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert len(results) == 2 and all(v < 1e-6 for v in results.values()), dict(results)
