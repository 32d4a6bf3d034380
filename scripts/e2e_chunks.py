"""e2e step time of the host-buffer API as a function of the number of (batch, head range) chunks."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fa2_triton_b200.host_api import HostAttention
B, S, H, D = 4, 4096, 32, 128
g = torch.Generator().manual_seed(0)
host = [torch.empty(B, S, H, D).normal_(0, 0.5, generator=g).bfloat16().pin_memory() for _ in range(4)]
for chunks in (4, 8, 16, 32, 64, 128):
    plan = HostAttention(host[0].shape, host[1].shape, torch.bfloat16, "cuda:0", chunks)
    for _ in range(2): plan(*host, causal=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): plan(*host, causal=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    print(f"chunks={plan.chunks:4d}: {dt * 1e3:6.2f} ms per step")
    del plan
