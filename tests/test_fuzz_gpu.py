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
