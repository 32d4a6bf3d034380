"""GPU: the BASELINE.json configurations at their full sizes.

The eager oracle materialises [B, H, Sq, Sk] fp32 (oracle/eager_reference.py), so at these sizes it checks sampled
(batch, head) slices - the kernels treat every slice alike - and the whole tensors are covered
by size-independent properties:
  * bit-repeatability of the deterministic backward over the full tensors (reference tests/test_repeatability.py),
  * exact linearity in dO for a power-of-two factor (every product, the delta row sum and each 16-bit rounding commute
    with a multiplication by 2),
  * V = 1  =>  O = 1 within the 16-bit tolerance and dQ = dK = 0 exactly (dP - delta vanishes when every dP row is
    constant ... up to rounding of O, so: tiny), which exercises masking, the row sums and the LSE at every row.
"""
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import ATOL_LSE, assert_close_16, generate_test_data
from oracle import eager_reference as er

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (name, B, H, S, D, causal): config 2/3, config 4 + its ragged tail cases, config 5 as sharded over 8 GPUs (B = 32 / 8)
CONFIGS = [
    ("cfg3", 4, 32, 4096, 128, True),
    ("cfg4", 2, 16, 16384, 64, False),
    ("cfg4-ragged", 2, 16, 1000, 64, False),
    ("cfg4-ragged-causal", 2, 16, 1000, 64, True),
    ("cfg5-shard", 4, 32, 8192, 128, True),
]


def run(q, k, v, do, causal):
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=causal)
    dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
    return out.detach(), dq, dk, dv


@pytest.mark.parametrize("name,B,H,S,D,causal", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_full_size_slices_match_oracle_and_properties_hold(name, B, H, S, D, causal):
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=3)
    out, dq, dk, dv = run(q, k, v, do, causal)
    o2, lse, _, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    assert torch.equal(o2, out)
    torch.cuda.synchronize()

    # sampled slices against the oracle (first, last and one inner (batch, head) pair)
    for b, h in {(0, 0), (B - 1, H - 1), (B // 2, H // 3)}:
        sl = lambda t: t[b : b + 1, :, h : h + 1].contiguous()
        ref = er.fwd_bwd(sl(q), sl(k), sl(v), sl(do), causal=causal, upcast=True)
        for nm, got, want in zip(("O", "dQ", "dK", "dV"), (out, dq, dk, dv), ref):
            assert_close_16(f"{name} {nm}[{b},{h}]", sl(got), want)
        lse_ref = er.lse_log2(sl(q), sl(k), causal=causal)
        torch.testing.assert_close(lse[b : b + 1, h : h + 1, :S], lse_ref, rtol=0, atol=ATOL_LSE)
        del ref, lse_ref
    torch.cuda.empty_cache()

    # bit-repeatability of the whole step (deterministic dQ)
    again = run(q, k, v, do, causal)
    for a, b_ in zip((out, dq, dk, dv), again):
        assert torch.equal(a, b_)

    # exact linearity in dO for a power of two
    _, dq2, dk2, dv2 = run(q, k, v, do * 2, causal)
    assert torch.equal(dq2, dq * 2) and torch.equal(dk2, dk * 2) and torch.equal(dv2, dv * 2)
    del again, dq2, dk2, dv2

    # V = 1: every output row is a convex combination of ones
    ones = torch.ones_like(v)
    o1, dq1, dk1, _ = run(q, k, ones, do, causal)
    assert_close_16(f"{name} O(V=1)", o1, torch.ones_like(o1))
    # dP[i, j] = sum_d dO[i, d] for every j and delta[i] = sum_d O[i, d] dO[i, d] with O ~ 1: dS = P (dP - delta) ~ 0
    bound = 4e-2 * do.float().abs().sum(-1).max().item()
    assert dq1.float().abs().max().item() <= bound and dk1.float().abs().max().item() <= bound


def test_full_size_atomic_mode_matches_deterministic():
    """Config 3 in the atomic fast mode: same result as the deterministic path within the 16-bit tolerance."""
    B, H, S, D = 4, 32, 4096, 128
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=5)
    _, dq, dk, dv = run(q, k, v, do, True)
    fa2.set_deterministic(False)
    try:
        _, dqa, dka, dva = run(q, k, v, do, True)
    finally:
        fa2.set_deterministic(True)
    assert_close_16("dQ", dqa, dq)
    assert_close_16("dK", dka, dk)
    assert_close_16("dV", dva, dv)


@pytest.mark.parametrize("name,B,H,S,D,causal", [c for c in CONFIGS if c[0] in ("cfg3", "cfg4", "cfg4-ragged-causal")],
                         ids=["cfg3", "cfg4", "cfg4-ragged-causal"])
def test_full_size_every_dq_mode_agrees(name, B, H, S, D, causal):
    """All dQ paths at full size: the deterministic ones (recomputing = 1, spilled dS = 5, ticket-ordered fused = 2) are
    bit-repeatable and agree with each other and with the atomic mode (0) within the 16-bit tolerance - the latent
    races of a long pipeline show up here (run-to-run differences), not at S = 1024."""
    q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=11)
    o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    ref = None
    for mode in (5, 1, 2, 0):
        a = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, causal, scale, None, deterministic=mode,
                                     dq_dtype=torch.bfloat16)
        if mode != 0:
            b = fa2._flash_attn_backward(do, q, k, v, None, None, o, lse, 0.0, causal, scale, None, deterministic=mode,
                                         dq_dtype=torch.bfloat16)
            for x, y in zip(a, b):
                assert torch.equal(x, y), f"mode {mode} is not repeatable"
            del b
        if ref is None:
            ref = a
        else:
            for nm, x, y in zip(("dQ", "dK", "dV"), a, ref):
                assert_close_16(f"{name} mode {mode} {nm}", x, y)
        torch.cuda.empty_cache()
