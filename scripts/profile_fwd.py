"""Small driver for ncu: a few forward launches (causal then non-causal), nothing else.
Usage: python scripts/profile_fwd.py [B H S D]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fa2_triton_b200 as fa2

a = sys.argv[1:]
B, H, S, D = (int(x) for x in (a[:4] if len(a) >= 4 else (4, 32, 4096, 128)))
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
for causal in (True, True, False, False):
    out = fa2.flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
