"""GPU parity of the forward kernel (through the C ABI via the Python host shim) against the oracle.

Mirrors the reference's tests/test_fwd_only.py + tests/core.py for the in-scope arguments, plus masking-exactness
probes and the golden vectors."""
import math

import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import (ATOL_LSE, assert_close_16, assert_reference_criterion, generate_test_data, golden_names,
                            load_golden)
from oracle import eager_reference as er

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_case(B, Sq, Sk, Hq, Hkv, D, causal, dtype, seed=0):
    q, k, v, _ = generate_test_data(B, Sq, Sk, Hq, Hkv, D, dtype, DEV, seed)
    o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    torch.cuda.synchronize()
    assert o.shape == q.shape and o.dtype == dtype
    assert lse.shape == (B, Hq, math.ceil(Sq / 128) * 128) and lse.dtype == torch.float32
    assert scale == pytest.approx(1 / math.sqrt(D))
    ref = er.flash_attn_reference(q, k, v, causal=causal, upcast=True)
    ref_lowp = er.flash_attn_reference(q, k, v, causal=causal, upcast=False, reorder_ops=True)
    assert_close_16("O", o, ref)
    assert_reference_criterion("O", o, ref, ref_lowp, 2, 5e-5)  # reference tests/utils.py:93
    lse_ref = er.lse_log2(q, k, causal=causal)
    got = lse[:, :, :Sq]
    finite = torch.isfinite(lse_ref)
    assert (torch.isfinite(got) == finite).all(), "LSE: -inf pattern differs"
    assert (got[finite] - lse_ref[finite]).abs().max().item() <= ATOL_LSE
    assert (lse[:, :, Sq:] == 0).all(), "LSE padding must be zero"
    return q, k, v, o, lse


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("Sq,Sk", [(128, 128), (256, 256), (512, 512), (1, 239), (113, 203), (128, 217), (108, 256),
                                   (256, 512), (1023, 1024), (1000, 1000), (384, 129)])
def test_forward_matches_oracle(Sq, Sk, D, causal, dtype):
    run_case(2, Sq, Sk, 3, 3, D, causal, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [136, 192, 207, 256])
@pytest.mark.parametrize("Sq,Sk", [(128, 128), (1, 239), (113, 203), (256, 512), (1023, 1024), (384, 129), (700, 64)])
def test_forward_head_dims_above_128(Sq, Sk, D, causal, dtype):
    """Head dims 129..256 (207 and 256 of the reference's grid, tests/test_fwd_bwd.py:16): the D = 256 instantiation of
    fwd2_sm100.cu - one Q tile per CTA, K / V tiles of 64 keys; 207 through a copy padded to 208."""
    run_case(2, Sq, Sk, 4, 2, D, causal, dtype, seed=D)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128, 256])
def test_forward_reference_maximum_moves_in_late_blocks(D, causal):
    """The reference maximum of a row only moves when a block exceeds it by 2^8 (lazy rescale), which random N(0, 0.5)
    inputs never do after the first block. Keys whose magnitude grows along the sequence make later blocks exceed the running
    maximum by far more than that, again and again: the O accumulator is rescaled in mid-row, and at head dim 64 the
    exponentials formed ahead of the block maximum are thrown away and redone (fwd2_sm100.cu). O and the LSE against the
    oracle, and run-to-run identical."""
    B, S, H = 2, 1024, 3
    q, k, v, _ = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, 11)
    ramp = torch.linspace(0.2, 6.0, S, device=DEV).view(1, S, 1, 1)
    k = (k.float() * ramp).to(torch.bfloat16)       # |scores| grow ~30x from the first to the last key
    q = (q.float() * 2.0).to(torch.bfloat16)
    o, lse, _, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    o2, lse2, _, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    assert torch.equal(o, o2) and torch.equal(lse, lse2)
    ref = er.flash_attn_reference(q, k, v, causal=causal, upcast=True)
    assert_close_16("O", o, ref)
    lse_ref = er.lse_log2(q, k, causal=causal)
    # the scores reach a few hundred in the log2 domain: compare the LSE relative to its size (fp32 has 24 bits)
    err = (lse[:, :, :S] - lse_ref).abs()
    assert (err <= ATOL_LSE + 2e-6 * lse_ref.abs()).all(), float(err.max())
    # the premise of the test: in most rows the maximum over the last quarter of the keys exceeds the maximum over the first
    # quarter by more than 2^8 in the log2 domain
    sc = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * (1.4426950408889634 / math.sqrt(D))
    if not causal:
        grows = (sc[..., 3 * S // 4:].amax(-1) - sc[..., : S // 4].amax(-1)) > 8.0
        assert grows.float().mean().item() > 0.5


def test_forward_long_rows_many_half_blocks():
    """Enough 64-key half-blocks for every ring / barrier phase of fwd2_sm100.cu to wrap several times, ragged ends,
    both tiles of a CTA with different block counts (causal) and Sq != Sk."""
    run_case(1, 1500, 2100, 2, 1, 128, True, torch.bfloat16, seed=3)
    run_case(1, 2100, 1500, 2, 2, 64, True, torch.bfloat16, seed=4)
    run_case(1, 1090, 2050, 2, 2, 256, False, torch.bfloat16, seed=5)


def test_forward_expanded_inputs_are_materialised():
    """k / v expanded over heads (stride 0) must give the same result as their materialised copies."""
    q, k, v, _ = generate_test_data(2, 300, 300, 4, 1, 64, torch.bfloat16, DEV, 11)
    ke, ve = k.expand(-1, -1, 4, -1), v.expand(-1, -1, 4, -1)
    assert ke.stride(2) == 0
    o1 = fa2.flash_attn_func(q, ke, ve, causal=True)
    o2 = fa2.flash_attn_func(q, ke.contiguous(), ve.contiguous(), causal=True)
    assert torch.equal(o1, o2)
    qb = q[:1].expand(2, -1, -1, -1)  # batch stride 0
    o3 = fa2.flash_attn_func(qb, k, v, causal=False)
    o4 = fa2.flash_attn_func(qb.contiguous(), k, v, causal=False)
    assert torch.equal(o3, o4)


@pytest.mark.parametrize("Hq,Hkv", [(8, 2), (4, 1), (9, 9)])
def test_forward_gqa(Hq, Hkv):
    run_case(2, 300, 300, Hq, Hkv, 128, True, torch.bfloat16)


def test_forward_fully_masked_rows():
    # causal, Sq > Sk: the first Sq - Sk queries see no key -> O = 0, LSE = -inf (src/forward/kernel.py:263-264)
    q, k, v, o, lse = run_case(1, 384, 129, 2, 2, 64, True, torch.bfloat16)
    assert (o[:, : 384 - 129] == 0).all()
    assert torch.isneginf(lse[:, :, : 384 - 129]).all()


def test_forward_strided_inputs():
    # q, k, v as slices of one packed projection: non-trivial batch/seq/head strides, last dim contiguous
    B, S, H, D = 2, 256, 4, 128
    g = torch.Generator(device="cpu").manual_seed(7)
    qkv = torch.empty(B, S, 3, H, D).normal_(0, 0.5, generator=g).to(torch.bfloat16).to(DEV)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o = fa2.flash_attn_func(q, k, v, causal=True)
    ref = er.flash_attn_reference(q, k, v, causal=True)
    assert_close_16("O", o, ref)


def test_forward_masking_is_exact():
    """Keys hidden by the causal mask or beyond Sk must have no influence at all: poison them and require
    bit-identical O and LSE."""
    B, Sq, Sk, H, D = 1, 200, 200, 2, 128
    q, k, v, _ = generate_test_data(B, Sq, Sk, H, H, D, torch.bfloat16, DEV, 3)
    o1, lse1, _, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)
    # with causal masking, query i never sees keys > i: changing key/value j only may change rows >= j
    k2, v2 = k.clone(), v.clone()
    k2[:, 150:] = 37.0
    v2[:, 150:] = -91.0
    o2, lse2, _, _ = fa2._flash_attn_forward(q, k2, v2, None, None, 0.0, True, None, None)
    assert torch.equal(o1[:, :150], o2[:, :150]) and torch.equal(lse1[:, :, :150], lse2[:, :, :150])
    # memory after the end of K/V (ragged tail inside a 128-row block) must be ignored as well
    kbig = torch.full((B, 256, H, D), 1e4, dtype=torch.bfloat16, device=DEV)
    vbig = torch.full((B, 256, H, D), 1e4, dtype=torch.bfloat16, device=DEV)
    kbig[:, :Sk], vbig[:, :Sk] = k, v
    o3, lse3, _, _ = fa2._flash_attn_forward(q, kbig[:, :Sk], vbig[:, :Sk], None, None, 0.0, False, None, None)
    o4, lse4, _, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, False, None, None)
    assert torch.equal(o3, o4) and torch.equal(lse3, lse4)


def test_forward_softmax_scale_argument():
    q, k, v, _ = generate_test_data(1, 128, 128, 2, 2, 64, torch.bfloat16, DEV, 5)
    o = fa2.flash_attn_func(q, k, v, softmax_scale=0.3)
    ref = er.flash_attn_reference(q, k, v, softmax_scale=0.3)
    assert_close_16("O", o, ref)


@pytest.mark.parametrize("name", [n for n in golden_names() if "d64" in n or "d128" in n])
def test_forward_matches_reference_golden(name):
    g = load_golden(name)
    q, k, v = (torch.from_numpy(g[n]).to(torch.bfloat16).to(DEV) for n in ("q", "k", "v"))
    o = fa2.flash_attn_func(q, k, v, causal=bool(g["causal"]))
    # golden inputs are fp32; rounding them to bf16 adds input error on top of the kernel's own
    assert_close_16("O", o, torch.from_numpy(g["out"]).to(DEV), atol=3e-2, rtol=2e-2)


def test_forward_repeatable():
    q, k, v, _ = generate_test_data(2, 512, 512, 4, 4, 128, torch.bfloat16, DEV, 9)
    outs = [fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)[:2] for _ in range(5)]
    for o, lse in outs[1:]:
        assert torch.equal(o, outs[0][0]) and torch.equal(lse, outs[0][1])


def test_race_condition_shapes():
    """The two shapes the reference singles out as once racy (tests/test_race_conditions.py:4-13), restated for
    the supported head dims: repeated runs must agree bit for bit."""
    for (B, H, Sq, Sk, D) in [(1, 1, 128, 128, 64), (4, 9, 127, 512, 64)]:
        q, k, v, _ = generate_test_data(B, Sq, Sk, H, H, D, torch.float16, DEV, 11)
        first = fa2.flash_attn_func(q, k, v)
        for _ in range(4):
            assert torch.equal(fa2.flash_attn_func(q, k, v), first)
