"""GPU parity of the reference's optional score modifiers - additive attention bias (forward and backward) and forward
dropout - against the oracle. Mirrors the `use_bias=True` / `dropout_p=0.1` arms of the reference's tests
(tests/test_fwd_bwd.py:16-20, tests/test_fwd_only.py:10-13, tests/core.py:28-29); the fixtures of the reference's own
Triton kernels for both are checked in tests/test_triton_golden_gpu.py."""
import numpy as np
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import ATOL_LSE, assert_close_16, assert_reference_criterion, generate_test_data
from oracle import eager_reference as er
from oracle import philox

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_bias(shape, dtype, seed, with_inf=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    bias = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1  # tests/core.py:28 uses rand(1, 1, Sq, Sk)
    if with_inf:  # a band mask expressed through the bias: every row keeps at least its first key
        Sq, Sk = shape[-2:]
        row = torch.arange(Sq).unsqueeze(-1)
        col = torch.arange(Sk)
        hide = (col > 0) & ((col - row * Sk // max(Sq, 1)).abs() > max(Sk // 3, 1))
        bias = bias.masked_fill(hide, float("-inf"))
    return bias.to(dtype).to(DEV)


def run_bias_case(B, Sq, Sk, Hq, Hkv, D, causal, dtype, bias_shape, bias_dtype=None, det=True, with_inf=False):
    q, k, v, do = generate_test_data(B, Sq, Sk, Hq, Hkv, D, dtype, DEV, seed=Sq + Sk)
    bias = make_bias(bias_shape, bias_dtype or dtype, seed=D, with_inf=with_inf)
    prev = fa2.is_deterministic()
    fa2.set_deterministic(det)
    try:
        qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, attention_bias=bias, causal=causal)
        dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
        _, lse, _, _ = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, causal, None, None)
        torch.cuda.synchronize()
    finally:
        fa2.set_deterministic(prev)
    ref = er.fwd_bwd(q, k, v, do, causal=causal, attn_bias=bias)
    ref_lowp = er.fwd_bwd(q, k, v, do, causal=causal, attn_bias=bias, upcast=False, reorder_ops=True)
    for name, got, want, low in zip(("O", "dQ", "dK", "dV"), (out, dq, dk, dv), ref, ref_lowp):
        assert_close_16(name, got, want)
        if not with_inf:
            assert_reference_criterion(name, got, want, low, 2 if name == "O" else 3, 5e-5 if name == "O" else 1e-5)
    lse_ref = er.lse_log2(q, k, causal=causal, attn_bias=bias)
    got = lse[:, :, :Sq]
    finite = torch.isfinite(lse_ref)
    assert (torch.isfinite(got) == finite).all()
    assert (got[finite] - lse_ref[finite]).abs().max().item() <= ATOL_LSE


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("Sq,Sk", [(1, 239), (127, 512), (113, 203), (128, 217), (108, 256), (256, 512), (384, 384)])
def test_bias_like_the_reference_tests(Sq, Sk, causal, dtype):
    # tests/core.py:28: one [1, 1, Sq, Sk] bias shared by every batch element and head
    run_bias_case(2, Sq, Sk, 4, 2, 64, causal, dtype, (1, 1, Sq, Sk))


@pytest.mark.parametrize("D", [128, 40])
@pytest.mark.parametrize("bias_shape", ["b1", "1h", "bh"])
def test_bias_broadcast_shapes(bias_shape, D):
    B, Sq, Sk, H = 3, 200, 328, 4
    shape = {"b1": (B, 1, Sq, Sk), "1h": (1, H, Sq, Sk), "bh": (B, H, Sq, Sk)}[bias_shape]
    run_bias_case(B, Sq, Sk, H, H, D, True, torch.bfloat16, shape)


@pytest.mark.parametrize("det", [True, False, 1, 5])
@pytest.mark.parametrize("D", [64, 128])
def test_bias_every_dq_mode(det, D):
    run_bias_case(2, 300, 300, 4, 2, D, True, torch.bfloat16, (1, 4, 300, 300), det=det)


def test_bias_fp32_and_strided():
    run_bias_case(2, 256, 256, 2, 2, 128, False, torch.bfloat16, (2, 2, 256, 256), bias_dtype=torch.float32)
    # a bias whose rows are not 16-byte aligned (odd Sk) and one that is a strided view of a larger tensor
    q, k, v, do = generate_test_data(1, 130, 333, 2, 2, 64, torch.float16, DEV, seed=3)
    big = make_bias((1, 2, 130, 400), torch.float16, seed=1)
    view = big[..., 5:338]
    out = fa2.flash_attn_func(q, k, v, attention_bias=view)
    assert_close_16("O", out, er.flash_attn_reference(q, k, v, attn_bias=view))


@pytest.mark.parametrize("dropout_p", [0.0, 0.1])
@pytest.mark.parametrize("B,H,Sq,Sk,D,causal,bias_bh", [
    (2, 4, 1024, 1280, 128, True, (1, 1)),    # many key blocks: every parity of the half-block barriers
    (2, 3, 700, 904, 64, False, (2, 3)),      # ragged tails in rows and keys, per-(batch, head) bias
    (1, 2, 130, 136, 128, True, (1, 2)),      # second Q tile of the pair nearly empty
    (2, 2, 1, 256, 64, False, (2, 1)),        # a single query row
])
def test_bias_staged_by_tma_equals_bias_through_plain_loads(B, H, Sq, Sk, D, causal, bias_bh, dropout_p):
    # A 16-bit bias TMA can address (16-byte aligned base and rows) is staged through shared memory
    # (fa2_fwd_kernel<.., kExt = 5 / 7>); any other takes the warps' own loads (kExt = 1 / 3). Same values, one element
    # of misalignment: O and LSE must not differ in a single bit.
    q, k, v, _ = generate_test_data(B, Sq, Sk, H, H, D, torch.bfloat16, DEV, seed=Sq + D)
    Bb, Hb = bias_bh
    aligned = make_bias((Bb, Hb, Sq, Sk), torch.bfloat16, seed=7)
    big = torch.zeros(Bb, Hb, Sq, Sk + 8, dtype=torch.bfloat16, device=DEV)
    view = big[..., 1:1 + Sk]
    view.copy_(aligned)
    assert aligned.data_ptr() % 16 == 0 and view.data_ptr() % 16 != 0
    a = fa2._flash_attn_forward(q, k, v, None, aligned, dropout_p, causal, None, 1234)
    b = fa2._flash_attn_forward(q, k, v, None, view, dropout_p, causal, None, 1234)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0])
    assert torch.equal(a[1], b[1])
    if dropout_p == 0.0:
        assert_close_16("O", a[0], er.flash_attn_reference(q, k, v, attn_bias=aligned, causal=causal))


@pytest.mark.parametrize("D,causal", [(128, True), (64, False)])
def test_bias_forward_is_bit_repeatable_over_many_runs(D, causal):
    # The TMA-staged bias buffer is overwritten while the kernel runs: a reader releasing it before its loads had
    # returned showed up as a few wrong rows in about 1 run of 100 at this size (scripts/bias_repeat.py). 300 runs.
    B, S, Hq, Hkv = 3, 640, 4, 2
    q, k, v, _ = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, 17)
    bias = make_bias((1, Hq, S, S), torch.bfloat16, seed=3)
    first = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, causal, None, None)
    for _ in range(300):
        again = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, causal, None, None)
        assert torch.equal(first[0], again[0]) and torch.equal(first[1], again[1])


def test_bias_minus_inf_acts_as_a_mask():
    run_bias_case(2, 256, 384, 2, 2, 128, False, torch.bfloat16, (1, 1, 256, 384), with_inf=True)
    run_bias_case(1, 200, 200, 2, 1, 64, True, torch.float16, (1, 2, 200, 200), with_inf=True, det=1)


def test_bias_only_some_gradients():
    q, k, v, do = generate_test_data(1, 256, 256, 2, 2, 128, torch.bfloat16, DEV, seed=5)
    bias = make_bias((1, 1, 256, 256), torch.bfloat16, seed=2)
    ref = er.fwd_bwd(q, k, v, do, causal=True, attn_bias=bias)
    qq = q.clone().requires_grad_(True)
    (dq,) = torch.autograd.grad(fa2.flash_attn_func(qq, k, v, attention_bias=bias, causal=True), (qq,), do)
    assert_close_16("dQ", dq, ref[1])
    kk, vv = k.clone().requires_grad_(True), v.clone().requires_grad_(True)
    dk, dv = torch.autograd.grad(fa2.flash_attn_func(q, kk, vv, attention_bias=bias, causal=True), (kk, vv), do)
    assert_close_16("dK", dk, ref[2])
    assert_close_16("dV", dv, ref[3])


def test_bias_argument_errors():
    q, k, v, _ = generate_test_data(2, 128, 128, 4, 4, 64, torch.bfloat16, DEV)
    with pytest.raises(ValueError):  # src/utils.py:67
        fa2.flash_attn_func(q, k, v, attention_bias=torch.zeros(3, 1, 128, 128, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(ValueError):  # src/utils.py:73
        fa2.flash_attn_func(q, k, v, attention_bias=torch.zeros(1, 3, 128, 128, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(AssertionError):  # src/utils.py:59
        fa2.flash_attn_func(q, k, v, attention_bias=torch.zeros(1, 1, 128, 100, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(AssertionError):  # src/forward/caller.py:27
        fa2.flash_attn_func(q, k, v, attention_mask=torch.ones(2, 128, dtype=torch.bool, device=DEV),
                            attention_bias=torch.zeros(1, 1, 128, 128, dtype=torch.bfloat16, device=DEV))


# ------------------------------------------------------ dropout ------------------------------------------------------
def keep_mask(seed, p, B, H, Sq, Sk):
    return torch.from_numpy(philox.dropout_keep_mask(seed, p, B, H, Sq, Sk)).to(DEV)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128, 59])
@pytest.mark.parametrize("Sq,Sk", [(1, 239), (127, 513), (128, 217), (256, 512)])
def test_dropout_forward_draws_the_reference_mask(Sq, Sk, D, causal, dtype):
    """tests/test_fwd_only.py:10 (dropout_p = 0.1): the oracle is given the keep-mask the reference's helper would
    generate for the seed (tests/utils.py:169-207) - equal outputs mean the kernel drew exactly that mask."""
    B, Hq, Hkv, p, seed = 2, 4, 2, 0.1, 123456789 + Sq
    q, k, v, _ = generate_test_data(B, Sq, Sk, Hq, Hkv, D, dtype, DEV, seed=Sq)
    o, lse, _, seed_out = fa2._flash_attn_forward(q, k, v, None, None, p, causal, None, seed)
    assert seed_out == seed
    keep = keep_mask(seed, p, B, Hq, Sq, Sk)
    ref = er.flash_attn_reference(q, k, v, causal=causal, dropout_p=p, dropout_mask=keep)
    assert_close_16("O", o, ref)
    # flipping a single mask bit moves O by more than the tolerance at these sizes: a tight check on top
    assert (o.float() - ref.float()).abs().max().item() <= 4e-3 + 8e-3 * ref.float().abs().max().item()
    lse_ref = er.lse_log2(q, k, causal=causal)  # the LSE is that of the undropped row
    finite = torch.isfinite(lse_ref)
    assert (lse[:, :, :Sq][finite] - lse_ref[finite]).abs().max().item() <= ATOL_LSE


def test_dropout_large_seed_bias_and_repeatability():
    B, Sq, Sk, H, D, p = 1, 200, 300, 2, 128, 0.5
    seed = (0xDEADBEEF << 32) | 0x12345678  # both key words in use
    q, k, v, _ = generate_test_data(B, Sq, Sk, H, H, D, torch.bfloat16, DEV, seed=1)
    bias = make_bias((1, 1, Sq, Sk), torch.bfloat16, seed=4)
    o1 = fa2.flash_attn_func(q, k, v, attention_bias=bias, dropout_p=p, dropout_seed=seed)
    o2 = fa2.flash_attn_func(q, k, v, attention_bias=bias, dropout_p=p, dropout_seed=seed)
    assert torch.equal(o1, o2)
    keep = keep_mask(seed, p, B, H, Sq, Sk)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    assert_close_16("O", o1, er.flash_attn_reference(q, k, v, attn_bias=bias, dropout_p=p, dropout_mask=keep))
    o3 = fa2.flash_attn_func(q, k, v, attention_bias=bias, dropout_p=p, dropout_seed=seed + 1)
    assert not torch.equal(o1, o3)
    # no seed given: one is drawn (src/utils.py:85)
    o4, _, _, drawn = fa2._flash_attn_forward(q, k, v, None, None, p, False, None, None)
    assert 0 <= drawn < 2**32 and torch.isfinite(o4).all()


@pytest.mark.parametrize("D", [64, 128, 59])
@pytest.mark.parametrize("causal", [False, True])
def test_dropout_with_attention_mask_numbers_its_counters_over_the_packed_batch(D, causal):
    """dropout together with attention_mask: the reference's VARLEN mode draws tl.rand at
    L * (start + L * head) + row * L + col for a sequence of length L starting at row `start` of the packed batch
    (src/forward/kernel.py:105-112, 147-149; pinned by tests/golden/triton_varlen_dropout_*.npz). The oracle gets that
    mask (oracle/philox.py); padded rows of O are zero; same seed -> same bits."""
    B, S, H = 4, 200, 3
    lengths = (200, 131, 1, 64)
    q, k, v, _ = generate_test_data(B, S, S, H, H, D, torch.float16, DEV, 5)
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    for b, n in enumerate(lengths):
        mask[b, :n] = True
    p, seed = 0.15, 2026
    o1 = fa2.flash_attn_func(q, k, v, attention_mask=mask, causal=causal, dropout_p=p, dropout_seed=seed)
    o2 = fa2.flash_attn_func(q, k, v, attention_mask=mask, causal=causal, dropout_p=p, dropout_seed=seed)
    assert torch.equal(o1, o2)
    keep = torch.from_numpy(philox.dropout_keep_mask_varlen(seed, p, lengths, H, S)).to(DEV)
    ref = er.flash_attn_reference(q, k, v, causal=causal, padding_mask=mask, dropout_p=p, dropout_mask=keep)
    assert_close_16("O", o1, ref)
    assert float(o1[~mask].abs().max()) == 0.0


def test_dropout_backward_is_not_implemented_like_the_reference():
    q, k, v, do = generate_test_data(1, 128, 128, 2, 2, 64, torch.bfloat16, DEV)
    qq = q.clone().requires_grad_(True)
    out = fa2.flash_attn_func(qq, k, v, dropout_p=0.1, dropout_seed=1)
    with pytest.raises(NotImplementedError):  # src/utils.py:87-88
        torch.autograd.grad(out, (qq,), do)


@pytest.mark.parametrize("D", [207, 256])
@pytest.mark.parametrize("causal", [False, True])
def test_bias_head_dims_above_128(D, causal):
    """The bias arm of the reference's grid at head dims 207 / 256 (tests/test_fwd_bwd.py:16-20): plain-load bias in
    fwd2_sm100.cu and bwd256_sm100.cu; 16-bit and fp32 biases, broadcast and per-(batch, head), -inf as a mask."""
    run_bias_case(2, 113, 203, 4, 2, D, causal, torch.bfloat16, (1, 1, 113, 203))
    run_bias_case(2, 256, 512, 4, 2, D, causal, torch.float16, (2, 4, 256, 512))
    run_bias_case(1, 200, 328, 2, 2, D, causal, torch.bfloat16, (1, 2, 200, 328), bias_dtype=torch.float32)
    run_bias_case(1, 256, 384, 2, 2, D, False, torch.bfloat16, (1, 1, 256, 384), with_inf=True)
