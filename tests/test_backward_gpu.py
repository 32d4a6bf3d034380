"""GPU parity of the backward path (delta + fused dK/dV/dQ + dQ convert, through the C ABI) against the oracle.
Mirrors the reference's tests/test_fwd_bwd.py / tests/core.py and states the intent of tests/test_repeatability.py."""
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import assert_close_16, assert_reference_criterion, generate_test_data, golden_names, load_golden
from oracle import eager_reference as er

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ours_fwd_bwd(q, k, v, do, causal, deterministic=True):
    fa2.set_deterministic(deterministic)
    try:
        q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(q, k, v, causal=causal)
        dq, dk, dv = torch.autograd.grad(out, (q, k, v), do)
    finally:
        fa2.set_deterministic(True)
    torch.cuda.synchronize()
    return out.detach(), dq, dk, dv


def run_case(B, Sq, Sk, Hq, Hkv, D, causal, dtype, deterministic=True, seed=0):
    q, k, v, do = generate_test_data(B, Sq, Sk, Hq, Hkv, D, dtype, DEV, seed)
    out, dq, dk, dv = ours_fwd_bwd(q, k, v, do, causal, deterministic)
    assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == v.shape
    assert dq.dtype == dtype and dk.dtype == dtype and dv.dtype == dtype
    ref = er.fwd_bwd(q, k, v, do, causal=causal, upcast=True)
    lowp = er.fwd_bwd(q, k, v, do, causal=causal, upcast=False, reorder_ops=True)
    for name, got, r, lp, mult, atol in (("O", out, ref[0], lowp[0], 2, 5e-5), ("dQ", dq, ref[1], lowp[1], 3, 1e-5),
                                         ("dK", dk, ref[2], lowp[2], 3, 1e-5), ("dV", dv, ref[3], lowp[3], 3, 1e-5)):
        assert_close_16(name, got, r)
        assert_reference_criterion(name, got, r, lp, mult, atol)  # reference tests/utils.py:93,126-140


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("Sq,Sk", [(128, 128), (256, 256), (512, 512), (1, 239), (113, 203), (128, 217), (108, 256),
                                   (256, 512), (1023, 1024), (1000, 1000), (384, 129)])
def test_backward_matches_oracle(Sq, Sk, D, causal, dtype):
    run_case(2, Sq, Sk, 3, 3, D, causal, dtype)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("mode", [0, 1, 2, 5])  # 0: atomic fast mode, 1: deterministic with a recomputing dQ kernel,
def test_backward_explicit_modes_match_oracle(D, causal, mode):  # 2: ticket-ordered fused kernel, 5: spilled dS
    run_case(2, 640, 640, 3, 3, D, causal, torch.bfloat16, deterministic=mode)
    run_case(1, 113, 203, 2, 2, D, causal, torch.bfloat16, deterministic=mode)
    run_case(1, 300, 1000, 4, 2, D, causal, torch.bfloat16, deterministic=mode)
    run_case(1, 1000, 300, 2, 1, D, causal, torch.bfloat16, deterministic=mode)


def test_backward_fp32_dq_like_the_reference():
    """`_flash_attn_backward` returns dq in fp32 by default (src/backward/caller.py:86), in every dQ mode."""
    q, k, v, do = generate_test_data(2, 300, 300, 4, 2, 128, torch.bfloat16, DEV, 8)
    o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    for mode in (0, 1, 2, 5, True):
        dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, None, deterministic=mode)
        assert dq.dtype == torch.float32 and dk.dtype == torch.bfloat16
        assert_close_16("dQ", dq, ref[1])
        assert_close_16("dK", dk, ref[2])
        assert_close_16("dV", dv, ref[3])


@pytest.mark.parametrize("Hq,Hkv", [(8, 2), (4, 1), (8, 1), (24, 3)])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_gqa(Hq, Hkv, causal):
    run_case(2, 300, 300, Hq, Hkv, 128, causal, torch.bfloat16)


@pytest.mark.parametrize("Hq,Hkv", [(32, 1), (32, 2)])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_large_head_groups(Hq, Hkv, causal):
    """16 and 32 query heads per kv head: dK / dV sum over the whole group (in fp32, inside one CTA) and reach 15-20,
    where one bf16 unit in the last place is 0.06-0.125 - the fixed 2e-2 + 1e-2 |ref| bound is then tighter than the
    output format (the same-precision eager arm misses it too, by more). The reference's own criterion applies: error
    at most 3 x the same-precision eager error (tests/utils.py:126-133)."""
    q, k, v, do = generate_test_data(2, 300, 300, Hq, Hkv, 128, torch.bfloat16, DEV, 0)
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=causal)
    grads = torch.autograd.grad(out, (qq, kk, vv), do)
    ref = er.fwd_bwd(q, k, v, do, causal=causal)
    lowp = er.fwd_bwd(q, k, v, do, causal=causal, upcast=False, reorder_ops=True)
    for name, got, r, lp in zip(("O", "dQ", "dK", "dV"), (out,) + tuple(grads), ref, lowp):
        assert_reference_criterion(name, got, r, lp, 2 if name == "O" else 3, 5e-5 if name == "O" else 1e-5)
        assert_close_16(name, got, r, atol=4e-2, rtol=1e-2)


@pytest.mark.parametrize("det", [False, 1, 5])
def test_backward_mqa_every_dq_mode_and_only_some_gradients(det):
    run_case(1, 384, 520, 16, 1, 64, True, torch.float16, deterministic=det)
    q, k, v, do = generate_test_data(2, 256, 256, 16, 2, 128, torch.bfloat16, DEV, 4)
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    kk, vv = k.clone().requires_grad_(True), v.clone().requires_grad_(True)
    dk, dv = torch.autograd.grad(fa2.flash_attn_func(q, kk, vv, causal=True), (kk, vv), do)
    assert dk.shape == k.shape and dv.shape == v.shape
    assert_close_16("dK", dk, ref[2])
    assert_close_16("dV", dv, ref[3])


def test_backward_fully_masked_rows_have_zero_dq():
    q, k, v, do = generate_test_data(1, 384, 129, 2, 2, 64, torch.bfloat16, DEV, 2)
    _, dq, dk, dv = ours_fwd_bwd(q, k, v, do, True)
    assert (dq[:, : 384 - 129] == 0).all()  # src/backward/compute_dq.py:248-249
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    assert_close_16("dK", dk, ref[2])
    assert_close_16("dV", dv, ref[3])


@pytest.mark.parametrize("mode", [True, 1, 2, 5])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
def test_deterministic_mode_is_bit_repeatable(D, causal, mode):
    """Intent of the reference's tests/test_repeatability.py:18-53 (10 repeats, identical results), checked on the
    full tensors instead of masked checksums."""
    q, k, v, do = generate_test_data(2, 1024, 1024, 4, 4, D, torch.bfloat16, DEV, 4)
    first = ours_fwd_bwd(q, k, v, do, causal, deterministic=mode)
    for _ in range(9):
        again = ours_fwd_bwd(q, k, v, do, causal, deterministic=mode)
        for a, b in zip(first, again):
            assert torch.equal(a, b)
    assert not any(torch.isnan(t).any() for t in first)


def test_spill_falls_back_to_recompute_when_the_workspace_is_capped(monkeypatch):
    """FA2_SPILL_MAX_GIB bounds the dS workspace; beyond it the deterministic path recomputes (mode 1) instead."""
    q, k, v, do = generate_test_data(1, 512, 512, 2, 2, 128, torch.bfloat16, DEV, 5)
    _, dq5, dk5, dv5 = ours_fwd_bwd(q, k, v, do, True, deterministic=True)
    monkeypatch.setenv("FA2_SPILL_MAX_GIB", "0")
    _, dq1, dk1, dv1 = ours_fwd_bwd(q, k, v, do, True, deterministic=True)
    _, dq1b, _, _ = ours_fwd_bwd(q, k, v, do, True, deterministic=1)
    assert torch.equal(dq1, dq1b)
    assert torch.equal(dk1, dk5) and torch.equal(dv1, dv5)
    assert_close_16("dQ", dq1, dq5.float())



def test_backward_masking_is_exact():
    """Causally hidden keys and the padding past Sk/Sq must not leak into any gradient."""
    B, S, H, D = 1, 200, 2, 128
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, 6)
    _, dq1, dk1, dv1 = ours_fwd_bwd(q, k, v, do, True)
    k2, v2 = k.clone(), v.clone()
    k2[:, 150:] = 37.0
    v2[:, 150:] = -91.0
    _, dq2, dk2, dv2 = ours_fwd_bwd(q, k2, v2, do, True)
    assert torch.equal(dq1[:, :150], dq2[:, :150])
    # gradients of the visible keys only receive contributions from queries >= key index; rows < 150 of dK/dV get
    # contributions from queries >= 150 too, which do see the poisoned keys -> only check the structure via dQ
    do2 = do.clone()
    do2[:, :100] = 0
    _, _, dk3, dv3 = ours_fwd_bwd(q, k, v, do2, True)
    _, _, dk4, dv4 = ours_fwd_bwd(q, k, v, do2, True)
    assert torch.equal(dk3, dk4) and torch.equal(dv3, dv4)


@pytest.mark.parametrize("name", [n for n in golden_names() if "d64" in n or "d128" in n])
def test_backward_matches_reference_golden(name):
    g = load_golden(name)
    q, k, v, do = (torch.from_numpy(g[n]).to(torch.bfloat16).to(DEV) for n in ("q", "k", "v", "do"))
    _, dq, dk, dv = ours_fwd_bwd(q, k, v, do, bool(g["causal"]))
    for nm, got in (("dq", dq), ("dk", dk), ("dv", dv)):
        assert_close_16(nm, got, torch.from_numpy(g[nm]).to(DEV), atol=6e-2, rtol=3e-2)


def test_delta_kernel():
    from fa2_triton_b200 import _lib
    lib = _lib.load()
    B, S, H, D = 2, 300, 3, 128
    o = torch.randn(B, S, H, D, device=DEV, dtype=torch.bfloat16)
    do = torch.randn(B, S, H, D, device=DEV, dtype=torch.bfloat16)
    delta = torch.full((B, H, 384), float("nan"), device=DEV)
    code = lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B, H, S, D, _lib.strides3(o),
                                   _lib.strides3(do), 384, _lib.DTYPE_BF16, None, 0,
                                   torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "delta")
    ref = (o.float() * do.float()).sum(-1).permute(0, 2, 1)
    torch.testing.assert_close(delta[:, :, :S], ref, rtol=1e-4, atol=1e-3)
    assert (delta[:, :, S:] == 0).all()


def test_backward_skips_unneeded_gradients():
    """Only q, or only k and v, require a gradient (ctx.needs_input_grad): the other kernel is not launched and the
    gradients that are produced equal those of the full backward."""
    q, k, v, do = generate_test_data(2, 384, 384, 4, 4, 128, torch.bfloat16, DEV, 12)
    _, dq_all, dk_all, dv_all = ours_fwd_bwd(q, k, v, do, True)
    qq = q.clone().requires_grad_(True)
    out = fa2.flash_attn_func(qq, k, v, causal=True)
    (dq_only,) = torch.autograd.grad(out, (qq,), do)
    # dQ alone comes from the recomputing kernel (nothing spills dS for it): same sum, exponentials evaluated on
    # different pipes -> equal within the 16-bit tolerance, not bit for bit
    assert_close_16("dQ", dq_only, dq_all.float())
    kk, vv = k.clone().requires_grad_(True), v.clone().requires_grad_(True)
    out = fa2.flash_attn_func(q, kk, vv, causal=True)
    dk_only, dv_only = torch.autograd.grad(out, (kk, vv), do)
    assert torch.equal(dk_only, dk_all) and torch.equal(dv_only, dv_all)


def test_kv_head_shards_equal_the_whole_call_bit_for_bit():
    """The batch x head partitioner cuts a call into (batch, kv-head) slices; the head-group split of the dK / dV kernel
    looks at the group size only, so every slice makes the whole call's choice: identical bits, GQA 16:2 (group 8, two
    parts per kv head) over 4 "ranks"."""
    from fa2_triton_b200 import shard
    B, S, Hq, Hkv, D = 2, 384, 16, 2, 128
    q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, 12)

    def fwd_bwd(qs, ks, vs, dos):
        qq, kk, vv = (t.detach().requires_grad_(True) for t in (qs, ks, vs))
        out = fa2.flash_attn_func(qq, kk, vv, causal=True)
        return (out.detach(),) + torch.autograd.grad(out, (qq, kk, vv), dos)

    whole = fwd_bwd(q, k, v, do)
    grp = Hq // Hkv
    for rank in range(4):
        for (b0, b1, h0, h1), qs, ks, vs, dos in shard.shard_views(q, k, v, 4, rank, do):
            part = fwd_bwd(qs, ks, vs, dos)
            assert torch.equal(part[0], whole[0][b0:b1, :, h0 * grp:h1 * grp])
            assert torch.equal(part[1], whole[1][b0:b1, :, h0 * grp:h1 * grp])
            assert torch.equal(part[2], whole[2][b0:b1, :, h0:h1])
            assert torch.equal(part[3], whole[3][b0:b1, :, h0:h1])


@pytest.mark.parametrize("kind", ["mqa_split", "varlen", "bias"])
def test_new_paths_are_bit_repeatable(kind):
    """tests/test_repeatability.py's intent (10 repeats, bit-identical outputs and gradients) for the paths added on top
    of the plain one: split head groups (fp32 partials summed in a fixed order), the in-kernel variable-length path and
    the bias variants."""
    B, S, D = 3, 640, 128
    Hq, Hkv = (16, 1) if kind == "mqa_split" else (4, 2)
    q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, 17)
    kw = {}
    if kind == "varlen":
        mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
        for b, n in enumerate((640, 301, 77)):
            mask[b, :n] = True
        kw["attention_mask"] = mask
    if kind == "bias":
        kw["attention_bias"] = torch.rand(1, Hq, S, S, device=DEV).to(torch.bfloat16)

    def once():
        qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, causal=True, **kw)
        return (out.detach(),) + torch.autograd.grad(out, (qq, kk, vv), do)

    first = once()
    for _ in range(9):
        for a, b in zip(first, once()):
            assert torch.equal(a, b)


def test_backward_expanded_dout_and_kv():
    """the dO autograd produces for out.sum() is an expanded tensor (every stride 0), and k / v may be
    expanded over heads; both must give the gradients of their materialised copies (the reference's pointer arithmetic
    takes such strides as they are)."""
    q, k, v, _ = generate_test_data(2, 300, 300, 4, 1, 128, torch.bfloat16, DEV, 12)

    def grads(expand):
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        ke, ve = (kk.expand(-1, -1, 4, -1), vv.expand(-1, -1, 4, -1)) if expand else (kk.repeat(1, 1, 4, 1), vv.repeat(1, 1, 4, 1))
        out = fa2.flash_attn_func(qq, ke, ve, causal=True)
        if expand:
            out.sum().backward()  # dO = ones, expanded
        else:
            out.backward(torch.ones_like(out))
        return qq.grad, kk.grad, vv.grad

    a, b = grads(True), grads(False)
    torch.cuda.synchronize()
    for name, x, y in zip(("dQ", "dK", "dV"), a, b):
        assert torch.equal(x, y), name
    ref = er.fwd_bwd(q, k.repeat(1, 1, 4, 1), v.repeat(1, 1, 4, 1), torch.ones_like(q), causal=True)
    assert_close_16("dQ", a[0], ref[1])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [136, 192, 207, 256])
@pytest.mark.parametrize("Sq,Sk", [(128, 128), (1, 239), (113, 203), (256, 512), (1023, 1024), (384, 129), (700, 64)])
def test_backward_head_dims_above_128(Sq, Sk, D, causal, dtype):
    """Head dims 129..256 (207 and 256 of the reference's grid, tests/test_fwd_bwd.py:16): bwd256_sm100.cu - dQ with
    spilled P / dS tiles, dK / dV from the tiles; GQA 4:2; 207 through copies padded to 208."""
    run_case(2, Sq, Sk, 4, 2, D, causal, dtype, seed=D)


def test_backward_head_dim_256_modes_and_repeatability():
    """fp32 dQ, skipped gradients and every `deterministic` request land on the one deterministic path; 5 runs bit-equal."""
    q, k, v, do = generate_test_data(2, 520, 650, 6, 2, 256, torch.bfloat16, DEV, 21)
    o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    base = None
    for det in (True, False, 1, 2, 5, True, True):
        dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, seed, deterministic=det)
        torch.cuda.synchronize()
        assert dq.dtype == torch.float32
        if base is None:
            base = (dq, dk, dv)
            for name, g, r in (("dQ", dq, ref[1]), ("dK", dk, ref[2]), ("dV", dv, ref[3])):
                assert_close_16(name, g, r)
        else:
            for name, a, b in zip(("dQ", "dK", "dV"), base, (dq, dk, dv)):
                assert torch.equal(a, b), f"{name} differs for deterministic={det}"
    dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, seed, need_dq=False)
    assert dq is None and torch.equal(dk, base[1]) and torch.equal(dv, base[2])
    dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, seed, need_dkdv=False)
    assert dk is None and dv is None and torch.equal(dq, base[0])
