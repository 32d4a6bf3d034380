"""Shared helpers for the parity tests (our restatement of the reference's tests/utils.py + tests/core.py)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# north-star tolerances (BASELINE.json): bf16 tensors |x - ref| <= 2e-2 + 1e-2 * |ref|, LSE abs 1e-3
ATOL_16, RTOL_16 = 2e-2, 1e-2
ATOL_LSE = 1e-3


def generate_test_data(B, Sq, Sk, Hq, Hkv, D, dtype, device, seed=0):
    """Same recipe as the reference's tests/utils.py:9-26: seed, Q/K/V ~ N(0, 0.5^2), dO ~ N(0, 1)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.empty(B, Sq, Hq, D, dtype=torch.float32).normal_(0, 0.5, generator=g)
    k = torch.empty(B, Sk, Hkv, D, dtype=torch.float32).normal_(0, 0.5, generator=g)
    v = torch.empty(B, Sk, Hkv, D, dtype=torch.float32).normal_(0, 0.5, generator=g)
    do = torch.empty(B, Sq, Hq, D, dtype=torch.float32).normal_(0, 1.0, generator=g)
    return tuple(t.to(dtype).to(device) for t in (q, k, v, do))


def assert_close_16(name, got, ref, atol=ATOL_16, rtol=RTOL_16):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    worst = (err - bound).max().item()
    assert worst <= 0, f"{name}: max |err| {err.max().item():.4e}, exceeds atol+rtol*|ref| by {worst:.4e}"


def assert_reference_criterion(name, got, ref_fp32, ref_lowp, mult, atol):
    """The reference's own acceptance test (tests/utils.py:93,126-140): our error against the fp32 oracle must be
    within `mult` x the error of the same-precision eager PyTorch computation (+ atol)."""
    e_ours = (got.float() - ref_fp32.float()).abs().max().item()
    e_pt = (ref_lowp.float() - ref_fp32.float()).abs().max().item()
    assert e_ours <= mult * e_pt + atol, f"{name}: ours {e_ours:.4e} > {mult} * pytorch {e_pt:.4e} + {atol}"


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def golden_names():
    """Fixtures from the reference's eager oracle (oracle/gen_golden.py)."""
    return sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith("triton_"))


def triton_golden_names(with_backward=None):
    """Fixtures from the reference's Triton kernels under Triton's CPU interpreter (oracle/gen_golden_triton.py).
    The forward-dropout cases hold no gradients (the reference's backward raises under dropout)."""
    names = sorted(p.stem for p in GOLDEN.glob("triton_*.npz"))
    if with_backward is None:
        return names
    return [n for n in names if ("dropout" not in n) == with_backward]


def triton_golden_extras(g):
    """(seqlens or None, bias or None, keep-mask or None, dropout_p) of a Triton fixture."""
    B, Sq, H, _ = g["q"].shape
    Sk = g["k"].shape[1]
    seqlens = g["attention_mask"].sum(axis=1) if "attention_mask" in g else None
    bias = g.get("bias")
    keep, p = None, 0.0
    if "dropout_keep" in g:
        keep = np.unpackbits(g["dropout_keep"])[: B * H * Sq * Sk].astype(bool).reshape(B, H, Sq, Sk)
        p = float(g["dropout_p"])
    elif "dropout_p" in g:
        # dropout together with attention_mask: no stored mask (the reference's helper cannot build it); the restated
        # packed-batch numbering (oracle/philox.py) is checked through O, which the Triton kernel produced
        from oracle import philox
        p = float(g["dropout_p"])
        keep = philox.dropout_keep_mask_varlen(int(g["dropout_seed"]), p, seqlens, H, Sq)
    return seqlens, bias, keep, p


def philox_keep_mask_torch(seed, p, B, H, Sq, Sk, device):
    """oracle/philox.py on the GPU (int64 arithmetic): the reference's dropout keep-mask for large shapes."""
    M32 = 0xFFFFFFFF
    c0 = torch.arange(B * H * Sq * Sk, dtype=torch.int64, device=device) & M32
    c1 = torch.zeros_like(c0)
    c2 = torch.zeros_like(c0)
    c3 = torch.zeros_like(c0)
    k0, k1 = int(seed) & M32, (int(seed) >> 32) & M32
    for _ in range(10):
        pa, pb = 0xD2511F53 * c0, 0xCD9E8D57 * c2  # < 2^64 as unsigned; int64 wraps, the low 32 bits and the
        hi_a = ((pa >> 32) & M32)                   # logical high word are recovered with masks
        hi_b = ((pb >> 32) & M32)
        c0, c1, c2, c3 = hi_b ^ c1 ^ k0, pb & M32, hi_a ^ c3 ^ k1, pa & M32
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    x = torch.where(c0 >= 2**31, c0 - 2**32, c0)  # reinterpret as int32
    x = torch.where(x < 0, -x - 1, x)
    u = x.to(torch.float32) * torch.tensor(4.6566127342e-10, dtype=torch.float32, device=device)
    return (u > torch.tensor(p, dtype=torch.float32, device=device)).reshape(B, H, Sq, Sk)
