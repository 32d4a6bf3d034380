#!/usr/bin/env python
"""Is the forward kernel power-bound? The same launches on random inputs and on ALL-ZERO inputs (tensor-core power is data
dependent: scripts/micro/umma_2cta.cu measured 2233 TFLOP/s with zero operands against 1789 with random ones), 300 launches
back to back each, with the SM clock and board power NVML reports during the loop."""
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fa2_triton_b200 as fa2  # noqa: E402

try:
    import pynvml
    pynvml.nvmlInit()
    H = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    pynvml = None

dev = torch.device("cuda", 0)
B, Hh, S, D = 4, 32, 4096, 128


def sample(stop, out):
    while not stop.is_set():
        if pynvml:
            out.append((pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0))
        time.sleep(0.003)


for causal in (False, True):
    for kind in ("random", "zeros", "random"):
        g = torch.Generator(device=dev).manual_seed(0)
        if kind == "random":
            q, k, v = (torch.empty(B, S, Hh, D, device=dev, dtype=torch.bfloat16).normal_(0, 0.5, generator=g) for _ in range(3))
        else:
            q, k, v = (torch.zeros(B, S, Hh, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
        for _ in range(5):
            fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
        torch.cuda.synchronize()
        stop, out = threading.Event(), []
        th = threading.Thread(target=sample, args=(stop, out), daemon=True)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300):
            fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / 300
        clk = sorted(c for c, _ in out)
        pw = sorted(p for _, p in out)
        med = lambda a: a[len(a) // 2] if a else None  # noqa: E731
        print(f"fwd B={B} H={Hh} S={S} D={D} causal={causal} inputs={kind}: {ms:.4f} ms/launch over 300 launches; "
              f"SM clock median {med(clk)} MHz (min {clk[0] if clk else None}), power median {med(pw)} W (max {pw[-1] if pw else None})",
              flush=True)
        time.sleep(1.0)
