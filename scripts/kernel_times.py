"""Times the individual kernels of one workload through the C ABI (CUDA events, L2-exceeding inputs).
Usage: [FA2_SM100_LIB=path] python scripts/kernel_times.py [B H S D causal reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math, torch
import fa2_triton_b200 as fa2
from fa2_triton_b200 import _lib

a = sys.argv[1:]
B, H, S, D = (int(x) for x in (a[:4] if len(a) >= 4 else (4, 32, 4096, 128)))
causal = bool(int(a[4])) if len(a) > 4 else True
reps = int(a[5]) if len(a) > 5 else 20
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
S_pad = math.ceil(S / 128) * 128
st = torch.cuda.current_stream().cuda_stream
delta = torch.empty_like(lse); dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
s3 = _lib.strides3
f_fwd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)

def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_fwd = timed(lambda: _lib.check(lib.fa2_sm100_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, H, S, S, D,
                                                   s3(q), s3(k), s3(v), s3(o), S_pad, scale, int(causal), 17, st), "fwd"))
out = {"fwd_ms": t_fwd, "fwd_tflops": f_fwd / t_fwd / 1e9}
import os
_only = os.environ.get("FA2_KT_MODES")  # e.g. "3,5": time only these backward modes
for mode, name in ((3, "dkdv"), (4, "dq"), (1, "det"), (5, "spill"), (6, "spill_dkdv"), (7, "spill_dq"), (0, "atomic")):
    if _only and str(mode) not in _only.split(","):
        continue
    wsb = lib.fa2_sm100_bwd_workspace_bytes(B, H, S, S, D, int(causal), mode)
    ws = torch.empty(max(wsb, 16), device=dev, dtype=torch.uint8)
    wp = ws.data_ptr() if wsb else None
    fused = mode in (0, 2)
    def bwd():
        _lib.check(lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B, H, S, D, s3(o), s3(do), S_pad, 17,
                                           wp if fused else None, wsb if fused else 0, st), "delta")
        _lib.check(lib.fa2_sm100_bwd(do.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                     dk.data_ptr(), dv.data_ptr(), wp, wsb, B, H, H, S, S, D, s3(do), s3(q), s3(k), s3(v), s3(dq), s3(dk), s3(dv),
                                     S_pad, scale, int(causal), mode, 17, 17, st), "bwd")
    t = timed(bwd)
    out[f"bwd_{name}_ms"] = t
    if mode in (0, 1, 5): out[f"bwd_{name}_tflops"] = 2.5 * f_fwd / t / 1e9
print({k_: round(v_, 4) for k_, v_ in out.items()})
