"""Per-kernel GPU durations (CUPTI via torch.profiler) of fwd + bwd in one dQ mode, averaged over a few steps.
Usage: python scripts/kernel_profile.py [mode=True|0|1|2|5] [B H S D causal steps]"""
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
import fa2_triton_b200 as fa2

a = sys.argv[1:]
mode = a[0] if a else "True"
mode = True if mode == "True" else int(mode)
B, H, S, D = (int(x) for x in (a[1:5] if len(a) >= 5 else (4, 32, 4096, 128)))
causal = bool(int(a[5])) if len(a) > 5 else True
steps = int(a[6]) if len(a) > 6 else 10
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16().requires_grad_(True) for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
fa2.set_deterministic(mode)

def step():
    out = fa2.flash_attn_func(q, k, v, causal=causal)
    torch.autograd.grad(out, (q, k, v), do)

for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps): step()
    torch.cuda.synchronize()
tot = defaultdict(float); cnt = defaultdict(int)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name] += e.device_time; cnt[e.name] += 1
print(f"mode={mode} B={B} H={H} S={S} D={D} causal={causal}: per-step kernel time (us), {steps} steps")
s = 0.0
for n, t in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"  {t / steps:9.1f}  x{cnt[n] / steps:.0f}  {n[:100]}")
    s += t / steps
print(f"  {s:9.1f}  total")
# idle time between consecutive kernels of the profiled region (launch gaps the CPU side is responsible for)
ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda x: x[0])
span = (ev[-1][1] - ev[0][0]) / steps
gaps = sum(max(0.0, b[0] - a[1]) for a, b in zip(ev, ev[1:])) / steps
print(f"  {span:9.1f}  wall per step (first kernel start to last kernel end), of which {gaps:.1f} us idle between kernels")
