"""Generate tests/golden/*.npz by running the REFERENCE's own CPU oracle (src/reference_implementation.py,
loaded by path so that its Triton-dependent package __init__ is bypassed) plus PyTorch autograd.

Run in the build container, where the reference is mounted at /root/reference (or $FA2_REFERENCE_DIR):
    python oracle/gen_golden.py
The .npz files are committed; nothing at test time needs the reference checkout.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path(os.environ.get("FA2_REFERENCE_DIR", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

# (name, B, Hq, Hkv, Sq, Sk, D, causal, dtype)
CASES = [
    ("config1_fp32_b2h4s128d64", 2, 4, 4, 128, 128, 64, False, torch.float32),  # BASELINE.json configs[0]
    ("causal_fp32_b1h2s128d64", 1, 2, 2, 128, 128, 64, True, torch.float32),
    ("causal_ragged_fp32_b1h2sq100sk100d64", 1, 2, 2, 100, 100, 64, True, torch.float32),
    ("causal_sq_lt_sk_fp32_b1h2sq64sk160d64", 1, 2, 2, 64, 160, 64, True, torch.float32),
    ("causal_sq_gt_sk_fp32_b1h2sq160sk64d64", 1, 2, 2, 160, 64, 64, True, torch.float32),
    ("gqa_fp32_b1h4kv2s96d128", 1, 4, 2, 96, 96, 128, False, torch.float32),
    ("bf16_causal_b1h2s256d128", 1, 2, 2, 256, 256, 128, True, torch.bfloat16),
]


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_impl", REF / "src" / "reference_implementation.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.flash_attn_reference


def main():
    if not REF.exists():
        sys.exit(f"reference checkout not found at {REF}")
    ref = load_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    for name, B, Hq, Hkv, Sq, Sk, D, causal, dtype in CASES:
        # same recipe as the reference's tests/utils.py:21-25 (seed 0; Q,K,V ~ N(0, 0.5), dO ~ N(0, 1))
        torch.manual_seed(0)
        q = torch.empty(B, Sq, Hq, D, dtype=dtype).normal_(0, 0.5)
        k = torch.empty(B, Sk, Hkv, D, dtype=dtype).normal_(0, 0.5)
        v = torch.empty(B, Sk, Hkv, D, dtype=dtype).normal_(0, 0.5)
        do = torch.randn_like(q)
        qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = ref(qr, kr, vr, causal=causal, upcast=True)
        dq, dk, dv = torch.autograd.grad(out, (qr, kr, vr), do)
        to_np = lambda t: t.detach().float().numpy()
        np.savez_compressed(
            OUT / f"{name}.npz",
            q=to_np(q), k=to_np(k), v=to_np(v), do=to_np(do),
            out=to_np(out), dq=to_np(dq), dk=to_np(dk), dv=to_np(dv),
            causal=np.array(causal), is_bf16=np.array(dtype == torch.bfloat16),
        )
        print(f"{name}: sum(out)={out.double().sum().item():.6f} sum(dq)={dq.double().sum().item():.6f} "
              f"sum(dv)={dv.double().sum().item():.6f}")


if __name__ == "__main__":
    main()
