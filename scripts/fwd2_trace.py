"""clock64() timeline of ONE CTA of the forward kernel (fwd2_sm100.cu) from the trace build of the library:

    make -C fa2_triton_b200/csrc TRACE=1
    FA2_SM100_LIB=fa2_triton_b200/libfa2_sm100_trace.so python scripts/fwd2_trace.py [causal 0|1 [B H S D]]

Per half-block j and tile t: how long the softmax thread waited for S (`wait`), its busy time split into load / max /
exp / store, and the round trip R(j) = [S(j+2) seen by the softmax thread] - [P(j) handed over]; for the issuer of the
tile how long it waited for P and how long issuing P V(j) + Q K^T(j+2) took.
"""
import ctypes
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("FA2_SM100_LIB", str(ROOT / "fa2_triton_b200" / "libfa2_sm100_trace.so"))
import fa2_triton_b200 as fa2  # noqa: E402
from fa2_triton_b200 import _lib  # noqa: E402

causal = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
lib = _lib.load()
dev = "cuda:0"
B, H, S, D = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (4, 32, 4096, 128)
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16() for _ in range(3))
for _ in range(3):
    fa2.flash_attn_func(q, k, v, causal=causal)
torch.cuda.synchronize()
n = 4 * 64 * 8
buf = (ctypes.c_longlong * n)()
fn = lib.fa2_sm100_debug_fetch_trace
fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_size_t]
fn.restype = ctypes.c_int
assert fn(0, buf, n) == 0
T = [[[buf[(r * 64 + it) * 8 + ev] for ev in range(8)] for it in range(64)] for r in range(4)]
t0 = min(x for r in T for it in r for x in it if x)
print(f"B={B} H={H} S={S} D={D} causal={causal}; cycles relative to the first stamp")
for t in (0, 1):
    sm, mm = T[t], T[2 + t]
    print(f"---- tile {t}: j | s_wait  load  max  exp  store | busy | P(j)->S(j+2) | issuer: p_wait issue | P(j) at")
    for j in range(64):
        if not sm[j][5]:
            continue
        e = sm[j]
        rt = (sm[j + 2][1] - e[5]) if j + 2 < 64 and sm[j + 2][1] else 0
        iw = (mm[j][1] - mm[j][0]) if mm[j][1] else 0
        ii = (mm[j][2] - mm[j][1]) if mm[j][2] else 0
        print(f"{j:3d} | {e[1] - e[0]:5d} {e[2] - e[1]:5d} {e[3] - e[2]:5d} {e[4] - e[3]:5d} {e[5] - e[4]:5d} | {e[5] - e[1]:5d} | "
              f"{rt:6d} | {iw:6d} {ii:6d} | {e[5] - t0:7d}")
