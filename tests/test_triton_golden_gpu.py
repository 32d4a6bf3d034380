"""GPU parity against the reference's OWN TRITON KERNELS: the fixtures tests/golden/triton_*.npz hold O, LSE, dQ, dK, dV
that `_fwd_kernel` / `_bwd_kernel` (src/forward/kernel.py, src/backward/kernel.py) produced for the stored fp16 inputs
under Triton's CPU interpreter (oracle/gen_golden_triton.py). Same inputs, same dtype, through the same entry points
(`_flash_attn_forward` / `_flash_attn_backward`, src/forward/caller.py:12, src/backward/caller.py:14).

Tolerance: the north star's |x - ref| <= 2e-2 + 1e-2 |ref| for O / dQ / dK / dV and 1e-3 for the LSE - and, tighter,
what two correct fp16 implementations with different tile sizes can differ by (a few fp16 ulps of the values)."""
import numpy as np
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import ATOL_LSE, assert_close_16, load_golden, triton_golden_extras, triton_golden_names

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(g):
    q, k, v, do = (torch.from_numpy(g[n]).to(DEV) for n in ("q", "k", "v", "do"))
    mask = torch.from_numpy(g["attention_mask"]).to(DEV) if "attention_mask" in g else None
    bias = torch.from_numpy(g["bias"]).to(DEV) if "bias" in g else None
    return q, k, v, do, mask, bias


@pytest.mark.parametrize("name", triton_golden_names())
def test_forward_matches_triton_kernel(name):
    g = load_golden(name)
    q, k, v, _, mask, bias = _inputs(g)
    seqlens, _, keep, p = triton_golden_extras(g)
    seed = int(g["dropout_seed"]) if keep is not None else None
    o, lse, scale, seed_out = fa2._flash_attn_forward(q, k, v, mask, bias, p, bool(g["causal"]), None, seed)
    torch.cuda.synchronize()
    assert scale == pytest.approx(float(g["softmax_scale"]))
    ref_o = torch.from_numpy(g["out"]).to(DEV)
    assert_close_16("O", o, ref_o)
    assert_close_16("O (tight)", o, ref_o, atol=2e-3, rtol=4e-3)
    Sq = q.shape[1]
    got = lse[:, :, :Sq].cpu().numpy()
    ref = g["lse"]
    rows = np.ones_like(ref, dtype=bool)
    if seqlens is not None:
        rows &= np.arange(Sq)[None, None, :] < seqlens[:, None, None]
    if bool(g["causal"]) and Sq > k.shape[1]:  # rows without a visible key: tile-size dependent in the reference
        rows[:, :, : Sq - k.shape[1]] = False
    assert np.abs(got[rows] - ref[rows]).max() <= ATOL_LSE


@pytest.mark.parametrize("deterministic", [True, False])
@pytest.mark.parametrize("name", triton_golden_names(with_backward=True))
def test_backward_matches_triton_kernel(name, deterministic):
    g = load_golden(name)
    q, k, v, do, mask, bias = _inputs(g)
    causal = bool(g["causal"])
    prev = fa2.is_deterministic()
    fa2.set_deterministic(deterministic)
    try:
        o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, mask, bias, 0.0, causal, None, None)
        # fp32 dQ like the reference's internal accumulator (src/backward/caller.py:86)
        dq, dk, dv = fa2._flash_attn_backward(do, q, k, v, bias, mask, o, lse, 0.0, causal, scale, seed,
                                              dq_dtype=torch.float32)
        torch.cuda.synchronize()
    finally:
        fa2.set_deterministic(prev)
    for name_, got, key in (("dQ", dq, "dq"), ("dK", dk, "dk"), ("dV", dv, "dv")):
        ref = torch.from_numpy(g[key]).to(DEV)
        assert_close_16(name_, got, ref)
        assert_close_16(name_ + " (tight)", got, ref, atol=4e-3, rtol=8e-3)


@pytest.mark.parametrize("name", [n for n in triton_golden_names(with_backward=True) if "bias" not in n])
def test_autograd_entry_point_matches_triton_kernel(name):
    g = load_golden(name)
    q, k, v, do, mask, _ = _inputs(g)
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, attention_mask=mask, causal=bool(g["causal"]))
    dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
    assert dq.dtype == q.dtype
    assert_close_16("O", out, torch.from_numpy(g["out"]).to(DEV))
    for name_, got, key in (("dQ", dq, "dq"), ("dK", dk, "dk"), ("dV", dv, "dv")):
        assert_close_16(name_, got, torch.from_numpy(g[key]).to(DEV))
