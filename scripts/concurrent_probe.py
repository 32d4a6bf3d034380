"""Experiment: how long do the dK/dV kernel (+ dS spill) and the dQ-from-spill kernel take when they run CONCURRENTLY on
disjoint sets of SMs (two streams, capped persistent grids) compared with back to back on all SMs? No dependency between
them here: the dQ kernel reads the tiles of an earlier call from a second workspace.
Usage: FA2_SM100_DKDV_PERSISTENT=1 FA2_SM100_DKDV_CTAS=108 FA2_SM100_DQ_CTAS=40 python scripts/concurrent_probe.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math, torch
import fa2_triton_b200 as fa2
from fa2_triton_b200 import _lib

B, H, S, D, causal = 4, 32, 4096, 128, True
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
S_pad = math.ceil(S / 128) * 128
delta = torch.empty_like(lse); dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
s3 = _lib.strides3
wsb = lib.fa2_sm100_bwd_workspace_bytes(B, H, S, S, D, 1, 5)
ws_a = torch.empty(wsb, device=dev, dtype=torch.uint8); ws_b = torch.empty(wsb, device=dev, dtype=torch.uint8)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

def bwd(mode, ws, stream):
    _lib.check(lib.fa2_sm100_bwd(do.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                 dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), wsb, B, H, H, S, S, D, s3(do), s3(q), s3(k), s3(v), s3(dq), s3(dk), s3(dv),
                                 S_pad, scale, int(causal), mode, 17, 17, stream.cuda_stream), "bwd")

st0 = torch.cuda.current_stream()
_lib.check(lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B, H, S, D, s3(o), s3(do), S_pad, 17, None, 0, st0.cuda_stream), "delta")
bwd(5, ws_a, st0); bwd(5, ws_b, st0)
torch.cuda.synchronize()
reps = 30
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st0)
    for _ in range(reps): fn()
    e1.record(st0); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def sequential():
    bwd(6, ws_a, st0); bwd(7, ws_b, st0)
def concurrent():
    sa.wait_stream(st0); sb.wait_stream(st0)
    bwd(6, ws_a, sa); bwd(7, ws_b, sb)
    st0.wait_stream(sa); st0.wait_stream(sb)
print({"dkdv_ctas": os.environ.get("FA2_SM100_DKDV_CTAS"), "dq_ctas": os.environ.get("FA2_SM100_DQ_CTAS"),
       "sequential_ms": round(timed(sequential), 4), "concurrent_ms": round(timed(concurrent), 4)})
