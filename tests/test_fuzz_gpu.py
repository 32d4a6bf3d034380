"""GPU: randomised shapes through the public API against the oracle - the reference's tests/test_fwd_bwd.py grid is a
fixed list; this draws from the same ranges (batch 1-4, heads with GQA ratios, ragged / unequal sequence lengths, head
dims incl. non-multiples of 8... of 16, both dtypes, causal or not, every dQ mode) with a fixed seed."""
import random

import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import assert_close_16, generate_test_data
from oracle import eager_reference as er

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cases(n, seed=20261018):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        hkv = rng.choice([1, 2, 3, 4])
        grp = rng.choice([1, 1, 2, 4])
        sq = rng.choice([1, 17, 64, 127, 128, 129, 200, 255, 256, 257, 384, 500, 777, 1024])
        sk = sq if rng.random() < 0.5 else rng.choice([1, 33, 128, 130, 256, 300, 511, 640, 1000])
        out.append(dict(B=rng.randint(1, 3), Sq=sq, Sk=sk, Hq=hkv * grp, Hkv=hkv, D=rng.choice([16, 32, 48, 64, 80, 96, 128]),
                        causal=rng.random() < 0.5, dtype=rng.choice([torch.bfloat16, torch.float16]),
                        mode=rng.choice([True, 0, 1, 2, 5]), seed=i))
    return out


@pytest.mark.parametrize("c", cases(40), ids=lambda c: f"B{c['B']}-Sq{c['Sq']}-Sk{c['Sk']}-H{c['Hq']}/{c['Hkv']}-D{c['D']}-"
                                                        f"{'c' if c['causal'] else 'n'}-{str(c['dtype'])[-4:]}-m{c['mode']}")
def test_random_shape_matches_oracle(c):
    q, k, v, do = generate_test_data(c["B"], c["Sq"], c["Sk"], c["Hq"], c["Hkv"], c["D"], c["dtype"], DEV, c["seed"])
    fa2.set_deterministic(c["mode"])
    try:
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, causal=c["causal"])
        dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
    finally:
        fa2.set_deterministic(True)
    ref = er.fwd_bwd(q, k, v, do, causal=c["causal"], upcast=True)
    for name, got, want in zip(("O", "dQ", "dK", "dV"), (out, dq, dk, dv), ref):
        assert got.shape == want.shape and got.dtype == c["dtype"]
        assert_close_16(name, got, want)


def extras_cases(n, seed=20261019):
    """Random draws over the optional arguments of the entry point: attention_mask (right-padded lengths incl. 0 and 1),
    attention_bias (every broadcast shape, 16-bit or fp32, sometimes with -inf entries), forward dropout."""
    rng = random.Random(seed)
    out = []
    for i in range(n):
        hkv = rng.choice([1, 2, 3])
        grp = rng.choice([1, 2, 4])
        kind = rng.choice(["mask", "mask", "bias", "bias", "dropout"])
        sq = rng.choice([1, 40, 127, 128, 129, 200, 256, 300, 513, 700])
        sk = sq if (kind == "mask" or rng.random() < 0.4) else rng.choice([1, 33, 128, 203, 256, 384, 640])
        out.append(dict(B=rng.randint(1, 4), Sq=sq, Sk=sk, Hq=hkv * grp, Hkv=hkv, D=rng.choice([8, 24, 40, 59, 64, 80, 111, 128]),
                        causal=rng.random() < 0.5, dtype=rng.choice([torch.bfloat16, torch.float16]), kind=kind,
                        mode=rng.choice([True, True, 0, 1, 5]), bshape=rng.choice(["11", "b1", "1h", "bh"]),
                        bias_fp32=rng.random() < 0.3, bias_inf=rng.random() < 0.3, p=rng.choice([0.05, 0.3]), seed=i))
    return out


@pytest.mark.parametrize("c", extras_cases(48), ids=lambda c: f"{c['kind']}-B{c['B']}-Sq{c['Sq']}-Sk{c['Sk']}-H{c['Hq']}/{c['Hkv']}-"
                                                               f"D{c['D']}-{'c' if c['causal'] else 'n'}-{str(c['dtype'])[-4:]}-m{c['mode']}")
def test_random_extras_match_oracle(c):
    from fa2_test_utils import philox_keep_mask_torch
    B, Sq, Sk, Hq = c["B"], c["Sq"], c["Sk"], c["Hq"]
    q, k, v, do = generate_test_data(B, Sq, Sk, Hq, c["Hkv"], c["D"], c["dtype"], DEV, 100 + c["seed"])
    g = torch.Generator(device="cpu").manual_seed(c["seed"])
    kw, ref_kw, fwd_only = {}, {}, False
    if c["kind"] == "mask":
        lengths = [int(torch.randint(0, Sq + 1, (1,), generator=g)) for _ in range(B)]
        lengths[int(torch.randint(0, B, (1,), generator=g))] = Sq
        mask = torch.zeros(B, Sq, dtype=torch.bool, device=DEV)
        for b, n in enumerate(lengths):
            mask[b, :n] = True
        kw, ref_kw = dict(attention_mask=mask), dict(padding_mask=mask)
    elif c["kind"] == "bias":
        shape = {"11": (1, 1), "b1": (B, 1), "1h": (1, Hq), "bh": (B, Hq)}[c["bshape"]] + (Sq, Sk)
        bias = torch.rand(shape, generator=g) * 4 - 2
        if c["bias_inf"]:  # hide a random 20 % of the entries but never a row's first key
            hide = torch.rand(shape, generator=g) < 0.2
            hide[..., 0] = False
            bias = bias.masked_fill(hide, float("-inf"))
        bias = bias.to(torch.float32 if c["bias_fp32"] else c["dtype"]).to(DEV)
        kw, ref_kw = dict(attention_bias=bias), dict(attn_bias=bias)
    else:
        seed = 1000003 * (c["seed"] + 1)
        kw = dict(dropout_p=c["p"], dropout_seed=seed)
        ref_kw = dict(dropout_p=c["p"], dropout_mask=philox_keep_mask_torch(seed, c["p"], B, Hq, Sq, Sk, DEV))
        fwd_only = True
    fa2.set_deterministic(c["mode"])
    try:
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, causal=c["causal"], **kw)
        grads = () if fwd_only else torch.autograd.grad(out, (qq, kk, vv), do)
    finally:
        fa2.set_deterministic(True)
    if fwd_only:
        assert_close_16("O", out, er.flash_attn_reference(q, k, v, causal=c["causal"], **ref_kw))
        return
    ref = er.fwd_bwd(q, k, v, do, causal=c["causal"], **ref_kw)
    for name, got, want in zip(("O", "dQ", "dK", "dV"), (out,) + tuple(grads), ref):
        assert got.shape == want.shape and got.dtype == c["dtype"]
        assert_close_16(name, got, want)
