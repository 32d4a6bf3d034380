"""Bit-repeatability of the bias paths (forward + backward 40 times, forward alone 400 times): how the race between a
thread's reads of the TMA-staged bias buffer and the next TMA write into it was found and its fix checked.
Usage: python scripts/bias_repeat.py"""
import sys, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import fa2_triton_b200 as fa2
from fa2_test_utils import generate_test_data
DEV = "cuda:0"
B, S, D, Hq, Hkv = 3, 640, 128, 4, 2
q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, 17)
bias = torch.rand(1, Hq, S, S, device=DEV).to(torch.bfloat16)
def once():
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=True, attention_bias=bias)
    return (out.detach(),) + torch.autograd.grad(out, (qq, kk, vv), do)
first = once()
bad = [0, 0, 0, 0]
for _ in range(40):
    for i, (a, b) in enumerate(zip(first, once())):
        if not torch.equal(a, b):
            bad[i] += 1
            if bad[i] == 1:
                d = (a.float() - b.float()).abs()
                idx = torch.nonzero(d > 0)
                print("first diff", i, "count", idx.shape[0], "max", d.max().item(), "where", idx[:5].tolist(), idx[-3:].tolist(), flush=True)
print("mismatches out/dq/dk/dv:", bad)
# forward alone
f0 = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, True, None, None)
nb = 0
for _ in range(400):
    f = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, True, None, None)
    nb += int(not (torch.equal(f[0], f0[0]) and torch.equal(f[1], f0[1])))
print("forward-only mismatches:", nb)
