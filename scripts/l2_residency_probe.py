"""Experiment: does the dQ-from-spill kernel find the dS tiles in L2 when it runs right after the dK/dV kernel that
wrote them, if they fit (H heads x 17 MB at S = 4096 against the 126 MB L2)? Times the dQ kernel (mode 7) directly
after mode 6 and after an L2 flush in between. Usage: python scripts/l2_residency_probe.py [H]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import math, torch
import fa2_triton_b200 as fa2
from fa2_triton_b200 import _lib
B, S, D, causal = 1, 4096, 128, True
H = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda:0"; lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
S_pad = math.ceil(S / 128) * 128
delta = torch.empty_like(lse); dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
s3 = _lib.strides3
wsb = lib.fa2_sm100_bwd_workspace_bytes(B, H, S, S, D, 1, 5)
ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream()
def bwd(mode):
    _lib.check(lib.fa2_sm100_bwd(do.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                 dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), wsb, B, H, H, S, S, D, s3(do), s3(q), s3(k), s3(v), s3(dq), s3(dk), s3(dv),
                                 S_pad, scale, int(causal), mode, 17, 17, st.cuda_stream), "bwd")
_lib.check(lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B, H, S, D, s3(o), s3(do), S_pad, 17, None, 0, st.cuda_stream), "delta")
res = {}
for name, do_flush in (("warm (dQ right after dK/dV)", False), ("cold (L2 flushed in between)", True)):
    ts = []
    for i in range(12):
        bwd(6)
        if do_flush: flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bwd(7); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    res[name] = round(sorted(ts)[len(ts) // 2], 1)
print(f"H={H}: dS = {wsb / 2**20:.0f} MiB; dQ kernel median us: {res}")
