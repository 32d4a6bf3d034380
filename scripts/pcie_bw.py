"""Pinned host <-> device copy bandwidth on this box: contiguous vs pitched (2 KB rows), one direction vs both."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fa2_triton_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
n = 512 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
def h2d_pitched():  # 16 slices of [4096 rows x 2 KB] out of 8 KB-pitch rows, x 4 tensors x 4 batch elements
    for t in range(16):
        base = t * (32 << 20)
        for g in range(4):
            lib.fa2_sm100_copy2d_async(d_in.data_ptr() + base + g * 2048, 8192, h_in.data_ptr() + base + g * 2048, 8192, 2048, 4096, 1, s1.cuda_stream)
for name, fn in (("H2D contiguous", h2d), ("D2H contiguous", d2h), ("both directions", both), ("H2D pitched 2 KB rows", h2d_pitched)):
    t = timed(fn)
    print(f"{name:24s} {t * 1e3:7.2f} ms  {n / t / 1e9:6.1f} GB/s per direction")
