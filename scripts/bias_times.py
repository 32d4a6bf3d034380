"""Forward / backward times with and without an attention bias (public entry points, CUDA events).
Usage: python scripts/bias_times.py [B H S D reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import fa2_triton_b200 as fa2

a = sys.argv[1:]
B, H, S, D = (int(x) for x in (a[:4] if len(a) >= 4 else (4, 32, 4096, 128)))
reps = int(a[4]) if len(a) > 4 else 10
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()


def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for label, bias in (("none", None), ("[1,1,S,S]", torch.rand(1, 1, S, S, device=dev).bfloat16()),
                    ("[1,H,S,S]", torch.rand(1, H, S, S, device=dev).bfloat16())):
    for causal in (True, False):
        o, lse, scale, seed = fa2._flash_attn_forward(q, k, v, None, bias, 0.0, causal, None, None)
        t_f = timed(lambda: fa2._flash_attn_forward(q, k, v, None, bias, 0.0, causal, None, None))
        t_b = timed(lambda: fa2._flash_attn_backward(do, q, k, v, bias, None, o, lse, 0.0, causal, scale, seed, dq_dtype=q.dtype))
        t_dkdv = timed(lambda: fa2._flash_attn_backward(do, q, k, v, bias, None, o, lse, 0.0, causal, scale, seed, need_dq=False))
        print(f"bias {label:10s} causal={causal}: fwd {t_f:.3f} ms, bwd (delta + dK/dV + dQ) {t_b:.3f} ms, delta + dK/dV only {t_dkdv:.3f} ms", flush=True)
