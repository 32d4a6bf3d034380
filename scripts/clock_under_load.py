"""SM clock / power while one kernel runs back to back for ~1.5 s each (NVML sampled every 5 ms)."""
import sys, threading, time, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, pynvml
import fa2_triton_b200 as fa2
from fa2_triton_b200 import _lib
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
B, H, S, D = 4, 32, 4096, 128
dev = "cuda:0"; lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, True, None, None)
delta = torch.empty_like(lse); dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
ws = torch.zeros(256, device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream; s3 = _lib.strides3; S_pad = S
lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B, H, S, D, s3(o), s3(do), S_pad, 17, None, 0, st)
def fwd(): lib.fa2_sm100_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, H, S, S, D, s3(q), s3(k), s3(v), s3(o), S_pad, scale, 1, 17, st)
def bwd(mode): lib.fa2_sm100_bwd(do.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), 256, B, H, H, S, S, D, s3(do), s3(q), s3(k), s3(v), s3(dq), s3(dk), s3(dv), S_pad, scale, 1, mode, 17, 17, st)
def mm():
    torch.matmul(A_, B_, out=C_)
A_ = torch.randn(8192, 8192, device=dev).bfloat16(); B_ = torch.randn(8192, 8192, device=dev).bfloat16(); C_ = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
for name, fn, flops in (("fwd", fwd, 0.5498e12), ("dkdv", lambda: bwd(3), 1.0996e12), ("dq", lambda: bwd(4), 0.2749e12), ("cublas 8192^3", mm, 2 * 8192.0**3)):
    samples = []; stop = False
    def sampler():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.005)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < 1.5:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop = True; th.join()
    ms = e0.elapsed_time(e1) / n
    cl = sorted(s_[0] for s_ in samples[len(samples)//3:]); pw = sorted(s_[1] for s_ in samples[len(samples)//3:])
    print(f"{name:14s} {ms:.4f} ms/launch  {flops/ms/1e9:8.1f} TFLOP/s  clock median {cl[len(cl)//2]} MHz (min {cl[0]}, max {cl[-1]})  power median {pw[len(pw)//2]:.0f} W")
