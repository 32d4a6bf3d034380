"""Effective SM clock while our kernels run: a fixed-cycle spin kernel (torch.cuda._sleep, clock64-based) is launched
between heavy kernels and timed with CUDA events; cycles / time = the SM frequency the power manager is granting.
Usage: python scripts/effective_clock.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import fa2_triton_b200 as fa2

dev = "cuda:0"
B, H, S, D = 4, 32, 4096, 128
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16().requires_grad_(True) for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
CYC = 400_000

def spin_mhz(n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): torch.cuda._sleep(CYC)
    e1.record(); torch.cuda.synchronize()
    return n * CYC / (e0.elapsed_time(e1) * 1e3)

def step():
    out = fa2.flash_attn_func(q, k, v, causal=True)
    torch.autograd.grad(out, (q, k, v), do)

time.sleep(1.0)
print(f"idle: spin kernel runs at {spin_mhz(5):.0f} MHz")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for name, fn in (("fwd+bwd", step), ("cublas 8192^3", lambda: a @ b)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); time.sleep(1.0)
    res = []
    evs = []
    for i in range(60):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.cuda._sleep(CYC); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    mhz = [CYC / (e0.elapsed_time(e1) * 1e3) for e0, e1 in evs]
    print(f"{name}: spin clock after step 1/2/3/5/10/20/40/60: " + " ".join(f"{mhz[i - 1]:.0f}" for i in (1, 2, 3, 5, 10, 20, 40, 60)) + " MHz")
