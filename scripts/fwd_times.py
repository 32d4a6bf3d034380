#!/usr/bin/env python
"""Forward kernel alone through the C ABI: CUDA events around back-to-back launches (20 = burst, 150 = sustained)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fa2_triton_b200 as fa2  # noqa: E402

dev = torch.device("cuda", 0)
CASES = [  # B, H, Hkv, Sq, Sk, D, causal
    (4, 32, 32, 4096, 4096, 128, True), (4, 32, 32, 4096, 4096, 128, False), (2, 16, 16, 16384, 16384, 64, False),
    (2, 16, 16, 16384, 16384, 64, True), (4, 32, 32, 4096, 4096, 256, True), (4, 32, 32, 4096, 4096, 96, True),
    (4, 32, 8, 4096, 4096, 128, True), (1, 32, 32, 1024, 8192, 128, True),
]
import os
for (B, H, Hkv, Sq, Sk, D, causal) in CASES[: int(os.environ.get('FWD_TIMES_CASES', '99'))]:
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.empty(B, Sq, H, D, device=dev, dtype=torch.bfloat16).normal_(0, 0.5, generator=g)
    k = torch.empty(B, Sk, Hkv, D, device=dev, dtype=torch.bfloat16).normal_(0, 0.5, generator=g)
    v = torch.empty(B, Sk, Hkv, D, device=dev, dtype=torch.bfloat16).normal_(0, 0.5, generator=g)
    vis = Sq * Sk if not causal else sum(min(Sk, max(0, i + Sk - Sq + 1)) for i in range(Sq))
    flops = 4.0 * B * H * vis * D
    out = {}
    for reps in (20, 150):
        for _ in range(3):
            fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
        e1.record()
        torch.cuda.synchronize()
        out[reps] = e0.elapsed_time(e1) / reps
    print(f"B={B} H={H}/{Hkv} Sq={Sq} Sk={Sk} D={D} causal={causal}: burst {out[20]:.4f} ms = {flops / out[20] / 1e9:.0f} TFLOP/s, "
          f"sustained {out[150]:.4f} ms = {flops / out[150] / 1e9:.0f} TFLOP/s", flush=True)
    del q, k, v
    torch.cuda.empty_cache()
