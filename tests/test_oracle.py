"""CPU tests: the numpy oracle and the eager-PyTorch port against the golden vectors produced by the reference's
own oracle (oracle/gen_golden.py), and against each other."""
import numpy as np
import pytest
import torch

from fa2_test_utils import golden_names, load_golden
from oracle import attention_oracle as ao
from oracle import eager_reference as er


@pytest.mark.parametrize("name", golden_names())
def test_numpy_oracle_matches_reference_golden(name):
    g = load_golden(name)
    causal = bool(g["causal"])
    o, _ = ao.attention_forward(g["q"], g["k"], g["v"], causal=causal)
    dq, dk, dv = ao.attention_backward(g["q"], g["k"], g["v"], g["do"], causal=causal)
    # golden tensors were computed in fp32 (and rounded to bf16 for the bf16 case)
    tol = dict(rtol=2e-2, atol=2e-2) if bool(g["is_bf16"]) else dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(o, g["out"], **tol)
    np.testing.assert_allclose(dq, g["dq"], **tol)
    np.testing.assert_allclose(dk, g["dk"], **tol)
    np.testing.assert_allclose(dv, g["dv"], **tol)


@pytest.mark.parametrize("name", golden_names())
def test_eager_port_matches_reference_golden(name):
    g = load_golden(name)
    dtype = torch.bfloat16 if bool(g["is_bf16"]) else torch.float32
    q, k, v, do = (torch.from_numpy(g[n]).to(dtype) for n in ("q", "k", "v", "do"))
    out, dq, dk, dv = er.fwd_bwd(q, k, v, do, causal=bool(g["causal"]))
    # same ops in the same order as the reference -> identical up to thread-count dependent summation order
    tol = dict(rtol=1e-2, atol=1e-2) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-6)
    for got, key in ((out, "out"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        np.testing.assert_allclose(got.float().numpy(), g[key], **tol)


def test_config1_checksums():
    """config-1 fp64 sums of the reference oracle's outputs at seed 0."""
    g = load_golden("config1_fp32_b2h4s128d64")
    assert abs(g["out"].astype(np.float64).sum() - (-82.419352)) < 1e-3
    assert abs(g["dq"].astype(np.float64).sum() - 3.158366) < 1e-3
    assert abs(g["dv"].astype(np.float64).sum() - (-194.486146)) < 1e-3
    o, _ = ao.attention_forward(g["q"], g["k"], g["v"], causal=False)
    assert abs(o.sum() - (-82.419352)) < 1e-3


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("sq,sk", [(64, 64), (37, 150), (150, 37), (1, 239)])
def test_lse_definitions_agree(causal, sq, sk):
    q, k, v, _ = ao.generate_test_data(2, 3, 3, sq, sk, 32, seed=1)
    _, lse_np = ao.attention_forward(q, k, v, causal=causal)
    lse_t = er.lse_log2(torch.from_numpy(q), torch.from_numpy(k), causal=causal).numpy()
    finite = np.isfinite(lse_np)
    assert (finite == np.isfinite(lse_t)).all()
    np.testing.assert_allclose(lse_np[finite], lse_t[finite], rtol=1e-5, atol=1e-5)
    # exp2(S*c - LSE) sums to 1 over visible keys: the property the backward kernel relies on
    s = np.einsum("bthd,bshd->bhts", q.astype(np.float64), k.astype(np.float64)) / np.sqrt(32) * ao.LOG2E
    vis = ao.visibility_mask(sq, sk, causal)
    p = np.where(vis[None, None], np.exp2(s - np.where(finite, lse_np, 0)[..., None]), 0.0)
    np.testing.assert_allclose(p.sum(-1)[finite], 1.0, rtol=1e-9)


def test_fully_masked_rows_are_zero():
    # causal with Sq > Sk: the first Sq-Sk queries see nothing (src/reference_implementation.py:107-108)
    q, k, v, do = ao.generate_test_data(1, 2, 2, 96, 32, 16, seed=3)
    o, lse = ao.attention_forward(q, k, v, causal=True)
    assert (o[:, :64] == 0).all() and np.isneginf(lse[..., :64]).all()
    dq, dk, dv = ao.attention_backward(q, k, v, do, causal=True)
    assert (dq[:, :64] == 0).all()
    out_t, dq_t, _, _ = er.fwd_bwd(*(torch.from_numpy(x) for x in (q, k, v, do)), causal=True)
    assert (out_t[:, :64] == 0).all() and (dq_t[:, :64] == 0).all()


def test_gqa_matches_expanded_heads():
    q, k, v, do = ao.generate_test_data(1, 4, 2, 40, 56, 16, seed=4)
    dq, dk, dv = ao.attention_backward(q, k, v, do, causal=True)
    kx, vx = np.repeat(k, 2, axis=2), np.repeat(v, 2, axis=2)
    dq2, dk2, dv2 = ao.attention_backward(q, kx, vx, do, causal=True)
    np.testing.assert_allclose(dq, dq2, rtol=1e-12)
    np.testing.assert_allclose(dk, dk2.reshape(1, 56, 2, 2, 16).sum(3), rtol=1e-12)
    np.testing.assert_allclose(dv, dv2.reshape(1, 56, 2, 2, 16).sum(3), rtol=1e-12)


def test_backward_is_gradient_of_forward():
    """Finite-difference check of the closed-form backward (independent of autograd)."""
    q, k, v, do = ao.generate_test_data(1, 1, 1, 6, 7, 4, dtype="float64", seed=5)
    dq, dk, dv = ao.attention_backward(q, k, v, do, causal=True)
    f = lambda q_, k_, v_: (ao.attention_forward(q_, k_, v_, causal=True)[0] * do).sum()
    eps = 1e-6
    for arr, grad in ((q, dq), (k, dk), (v, dv)):
        idx = (0, 2, 0, 1)
        hi = arr.copy(); hi[idx] += eps
        lo = arr.copy(); lo[idx] -= eps
        args_hi = [hi if a is arr else a for a in (q, k, v)]
        args_lo = [lo if a is arr else a for a in (q, k, v)]
        num = (f(*args_hi) - f(*args_lo)) / (2 * eps)
        assert abs(num - grad[idx]) < 1e-6


# ------------------------------------------------------------------------------------------------------------------
# Fixtures computed by the reference's own Triton kernels (Triton's CPU interpreter, oracle/gen_golden_triton.py)
# ------------------------------------------------------------------------------------------------------------------
from fa2_test_utils import triton_golden_extras, triton_golden_names  # noqa: E402
from oracle import philox  # noqa: E402


@pytest.mark.parametrize("name", triton_golden_names())
def test_numpy_oracle_matches_triton_kernel_forward(name):
    """O and the log2-domain LSE of the reference's `_fwd_kernel` (fp16 inputs, fp32 accumulation) against the float64
    restatement: the LSE within the north star's 1e-3, O within fp16 output rounding + fp16 P rounding."""
    g = load_golden(name)
    seqlens, bias, keep, p = triton_golden_extras(g)
    o, lse = ao.attention_forward(g["q"], g["k"], g["v"], causal=bool(g["causal"]), bias=bias, seqlens=seqlens,
                                  dropout_keep=keep, dropout_p=p)
    np.testing.assert_allclose(o, g["out"].astype(np.float64), rtol=1e-2, atol=2e-3)
    finite = np.isfinite(lse)
    if seqlens is not None:  # rows of padded queries: the reference leaves whatever its buffer held
        finite &= (np.arange(lse.shape[-1])[None, None, :] < seqlens[:, None, None])
    ref_lse = g["lse"].astype(np.float64)
    # rows without any visible key (causal, Sq > Sk) are not compared: the reference leaves its zero-initialised buffer
    # untouched where a whole tile is hidden (src/forward/kernel.py:128-130) and stores -inf / NaN arithmetic where only
    # part of one is - a tile-size dependent value; the oracle (and the CUDA kernel) say -inf. O = 0 there for all.
    np.testing.assert_allclose(lse[finite], ref_lse[finite], rtol=0, atol=1e-3)


@pytest.mark.parametrize("name", triton_golden_names(with_backward=True))
def test_numpy_oracle_matches_triton_kernel_backward(name):
    g = load_golden(name)
    seqlens, bias, _, _ = triton_golden_extras(g)
    dq, dk, dv = ao.attention_backward(g["q"], g["k"], g["v"], g["do"], causal=bool(g["causal"]), bias=bias,
                                       seqlens=seqlens)
    # the kernel rounds P and dS to fp16 before the products and dK / dV to fp16 at the end; dQ stays fp32
    np.testing.assert_allclose(dq, g["dq"].astype(np.float64), rtol=1e-2, atol=3e-3)
    np.testing.assert_allclose(dk, g["dk"].astype(np.float64), rtol=1e-2, atol=4e-3)
    np.testing.assert_allclose(dv, g["dv"].astype(np.float64), rtol=1e-2, atol=4e-3)


@pytest.mark.parametrize("name", triton_golden_names())
def test_eager_port_matches_triton_kernel(name):
    g = load_golden(name)
    seqlens, bias, keep, p = triton_golden_extras(g)
    q, k, v, do = (torch.from_numpy(g[n]) for n in ("q", "k", "v", "do"))
    kw = dict(causal=bool(g["causal"]),
              padding_mask=None if seqlens is None else torch.from_numpy(g["attention_mask"]),
              attn_bias=None if bias is None else torch.from_numpy(bias))
    if keep is not None:
        out = er.flash_attn_reference(q, k, v, dropout_p=p, dropout_mask=torch.from_numpy(keep), **kw)
        np.testing.assert_allclose(out.float().numpy(), g["out"].astype(np.float32), rtol=1e-2, atol=2e-3)
        return
    out, dq, dk, dv = er.fwd_bwd(q, k, v, do, **kw)
    for got, key, atol in ((out, "out", 2e-3), (dq, "dq", 3e-3), (dk, "dk", 4e-3), (dv, "dv", 4e-3)):
        np.testing.assert_allclose(got.float().numpy(), g[key].astype(np.float32), rtol=1e-2, atol=atol)


@pytest.mark.parametrize("name", [n for n in triton_golden_names(with_backward=False) if "varlen" not in n])
def test_philox_restatement_is_bit_exact(name):
    """tl.rand(seed, linear offset) > p, as drawn by the reference (tests/utils.py:169-207): every bit of the mask."""
    g = load_golden(name)
    B, Sq, H, _ = g["q"].shape
    _, _, keep, p = triton_golden_extras(g)
    mine = philox.dropout_keep_mask(int(g["dropout_seed"]), p, B, H, Sq, g["k"].shape[1])
    assert (mine == keep).all()
    assert abs(keep.mean() - (1 - p)) < 0.01


def test_philox_known_answer():
    """Philox-4x32-10 known-answer vectors (Random123 kat_vectors: zero counter / zero key; all-ones)."""
    assert philox.philox4x32_10_first_word(0, np.array([0]))[0] == 0x6627E8D5
    # counter (ffffffff,0,0,0) differs from the all-ones KAT, so only check the stream is a function of seed and offset
    a = philox.philox4x32_10_first_word(123, np.arange(8))
    b = philox.philox4x32_10_first_word(123, np.arange(8))
    c = philox.philox4x32_10_first_word(124, np.arange(8))
    assert (a == b).all() and (a != c).any() and len(set(a.tolist())) == 8
