"""Small driver for ncu: a few fwd+bwd steps of one workload through the public API, nothing else.
Usage: python scripts/profile_target.py [B H S D causal steps [bias]]   (bias: a [1, 1, S, S] attention bias)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fa2_triton_b200 as fa2

a = sys.argv[1:]
B, H, S, D = (int(x) for x in (a[:4] if len(a) >= 4 else (4, 32, 4096, 128)))
causal = bool(int(a[4])) if len(a) > 4 else True
steps = int(a[5]) if len(a) > 5 else 3
with_bias = len(a) > 6 and a[6] == "bias"
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
kw = {"attention_bias": torch.rand(1, 1, S, S, device=dev, generator=g).bfloat16()} if with_bias else {}
for _ in range(steps):
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=causal, **kw)
    grads = torch.autograd.grad(out, (qq, kk, vv), do)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(grads[0].float().abs().mean()))
