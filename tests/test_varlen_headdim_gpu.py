"""GPU: the "next" rows of SURVEY.md section 8f that live on the host side of the boundary -
variable-length batches through `attention_mask` (reference tests/core.py with use_attention=True) and head dims
below 128 that are not 64 / 128 (zero-padded; reference tests/test_fwd_bwd.py:16-35 head_dim grid)."""
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import assert_close_16, assert_reference_criterion, generate_test_data
from oracle import eager_reference as er

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def right_padded_mask(B, S, seed):
    """The reference's generate_attention_mask (tests/utils.py:40-56): random right padding, one full sequence."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    pad = torch.randint(0, S - 1, (B,), generator=g).tolist()
    pad[int(torch.randint(0, B, (1,), generator=g))] = 0
    m = torch.ones(B, S, dtype=torch.bool)
    for i, p in enumerate(pad):
        if p:
            m[i, -p:] = False
    return m.to(DEV)


def ours(q, k, v, do, causal, mask=None):
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, attention_mask=mask, causal=causal)
    dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
    torch.cuda.synchronize()
    return out.detach(), dq, dk, dv


def check(got, q, k, v, do, causal, mask=None):
    ref = er.fwd_bwd(q, k, v, do, causal=causal, upcast=True, padding_mask=mask)
    lowp = er.fwd_bwd(q, k, v, do, causal=causal, upcast=False, reorder_ops=True, padding_mask=mask)
    for name, g, r, lp, mult, atol in (("O", got[0], ref[0], lowp[0], 2, 5e-5), ("dQ", got[1], ref[1], lowp[1], 3, 1e-5),
                                       ("dK", got[2], ref[2], lowp[2], 3, 1e-5), ("dV", got[3], ref[3], lowp[3], 3, 1e-5)):
        assert g.shape == r.shape and g.dtype == q.dtype
        assert_close_16(name, g, r)
        assert_reference_criterion(name, g, r, lp, mult, atol)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("B,S,Hq,Hkv", [(3, 300, 4, 2), (4, 1024, 2, 2), (2, 113, 3, 3)])
def test_attention_mask_matches_oracle(B, S, Hq, Hkv, D, causal):
    q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, seed=B + S)
    mask = right_padded_mask(B, S, seed=S)
    got = ours(q, k, v, do, causal, mask)
    check(got, q, k, v, do, causal, mask)
    # padded positions: exactly zero outputs and gradients (src/utils.py:26 unpacks into zeros)
    for t in got:
        assert (t[~mask] == 0).all()


def test_attention_mask_ignores_padded_content():
    """Whatever sits in the padded rows of q, k, v and dO must not reach any result."""
    B, S, H, D = 3, 384, 2, 128
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=9)
    mask = right_padded_mask(B, S, seed=4)
    a = ours(q, k, v, do, True, mask)
    q2, k2, v2, do2 = (t.clone() for t in (q, k, v, do))
    for t, val in ((q2, 31.0), (k2, -17.0), (v2, 55.0), (do2, 3.0)):
        t[~mask] = val
    b = ours(q2, k2, v2, do2, True, mask)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_attention_mask_padding_may_hold_nans_and_lengths_zero_and_one():
    """The per-sequence tensor maps end at each sequence's length, so TMA never reads the padding: NaNs there cannot
    reach a result (0 x NaN in a matrix product would). Lengths 0 and 1 are legal."""
    B, S, H, D = 4, 300, 2, 128
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=21)
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    for b, n in enumerate((300, 129, 1, 0)):
        mask[b, :n] = True
    a = ours(q, k, v, do, True, mask)
    check(a, q, k, v, do, True, mask)
    q2, k2, v2, do2 = (t.clone() for t in (q, k, v, do))
    for t in (q2, k2, v2, do2):
        t[~mask] = float("nan")
    b_ = ours(q2, k2, v2, do2, True, mask)
    for x, y in zip(a, b_):
        assert torch.isfinite(y).all() and torch.equal(x, y)


@pytest.mark.parametrize("D", [40, 64, 96])
def test_attention_mask_one_launch_equals_one_launch_per_sequence(D, monkeypatch):
    """The in-kernel variable-length path against the per-sequence path (FA2_VARLEN_PER_SEQUENCE=1): same kernels, same
    tiles, same order of operations - bit-identical results, fp32 dQ included."""
    B, S, Hq, Hkv = 3, 520, 4, 2
    q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.float16, DEV, seed=D)
    mask = right_padded_mask(B, S, seed=D)

    def run():
        o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, mask, None, 0.0, True, None, None)
        dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, mask, o, lse, 0.0, True, scale, seed, deterministic=5)
        torch.cuda.synchronize()
        return o, lse, dq, dk, dv

    one = run()
    monkeypatch.setenv("FA2_VARLEN_PER_SEQUENCE", "1")
    per_seq = run()
    for name, x, y in zip(("O", "LSE", "dQ", "dK", "dV"), one, per_seq):
        assert torch.equal(x, y), name


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [8, 16, 24, 32, 40, 59, 72, 80, 96, 104, 111, 120])
def test_other_head_dims_match_oracle(D, causal):
    """Multiples of 8 go to the kernels as they are (TMA zero-fills up to the kernel's 64 / 128 columns), the others
    (59, 111 of the reference's grid, tests/test_fwd_only.py:12) through zero-padded copies."""
    B, Sq, Sk, Hq, Hkv = 2, 200, 328, 4, 2
    q, k, v, do = generate_test_data(B, Sq, Sk, Hq, Hkv, D, torch.bfloat16, DEV, seed=D)
    assert q.stride(-1) == 1
    check(ours(q, k, v, do, causal), q, k, v, do, causal)


@pytest.mark.parametrize("det", [False, 1, 5, 2])
@pytest.mark.parametrize("D", [40, 96])
def test_other_head_dims_every_dq_mode(D, det):
    """fp32 dQ (the reference's convention, src/backward/caller.py:86) and every way of summing dQ; the fused modes
    fall back to padded copies, the deterministic ones run on the tensors as they are."""
    q, k, v, do = generate_test_data(2, 300, 300, 4, 2, D, torch.float16, DEV, seed=D)
    prev = fa2.is_deterministic()
    fa2.set_deterministic(det)
    try:
        o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)
        dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, seed)
        torch.cuda.synchronize()
    finally:
        fa2.set_deterministic(prev)
    assert dq.dtype == torch.float32 and dq.shape == q.shape and dk.shape == k.shape
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    for name, g, r in (("O", o, ref[0]), ("dQ", dq, ref[1]), ("dK", dk, ref[2]), ("dV", dv, ref[3])):
        assert_close_16(name, g, r)


def test_native_head_dim_strided_views_and_no_stray_writes():
    """q, k, v as slices of one packed [B, S, 3, H, D] tensor with D = 80, outputs written into slices of larger
    buffers: nothing outside the D columns may be touched (TMA clips the store at the tensor's head dim)."""
    B, S, H, D = 2, 260, 3, 80
    qkv = torch.randn(B, S, 3, H, D, device=DEV, dtype=torch.bfloat16) * 0.5
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    do = torch.randn(B, S, H, D, device=DEV, dtype=torch.bfloat16)
    big = torch.full((B, S, H, 128), 7.0, device=DEV, dtype=torch.bfloat16)
    o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None, out=big[..., :D])
    assert (big[..., D:] == 7.0).all()
    dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, True, scale, seed, dq_dtype=q.dtype)
    ref = er.fwd_bwd(q, k, v, do, causal=True)
    for name, g, r in (("O", o, ref[0]), ("dQ", dq, ref[1]), ("dK", dk, ref[2]), ("dV", dv, ref[3])):
        assert_close_16(name, g, r)


def test_head_dim_above_256_is_rejected():
    q, k, v, _ = generate_test_data(1, 128, 128, 2, 2, 264, torch.bfloat16, DEV)
    with pytest.raises(NotImplementedError):
        fa2.flash_attn_func(q, k, v)


def test_varlen_abi_clamps_out_of_range_lengths():
    """fa2_sm100_fwd_varlen with lengths outside [0, S] (only reachable through the C ABI): treated as S and 0."""
    import math
    from fa2_triton_b200 import _lib
    lib = _lib.load()
    B, S, H, D = 3, 200, 2, 64
    q, k, v, _ = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=31)

    def run(lengths):
        o = torch.zeros_like(q)
        S_pad = math.ceil(S / 128) * 128
        lse = torch.zeros(B, H, S_pad, device=DEV, dtype=torch.float32)
        sl = torch.tensor(lengths, device=DEV, dtype=torch.int32)
        ws = torch.empty(lib.fa2_sm100_varlen_workspace_bytes(B), device=DEV, dtype=torch.uint8)
        rc = lib.fa2_sm100_fwd_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, H, S, D,
                                      _lib.strides3(q), _lib.strides3(k), _lib.strides3(v), _lib.strides3(o), S_pad,
                                      1.0 / math.sqrt(D), 1, 17, sl.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fa2_sm100_fwd_varlen")
        torch.cuda.synchronize()
        return o, lse

    a = run([S + 57, -4, 77])
    b = run([S, 0, 77])
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert (a[0][1] == 0).all()
