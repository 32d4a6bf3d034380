"""clock64() pipeline timeline of ONE CTA of each kernel, from the trace build of the library.

    make -C fa2_triton_b200/csrc TRACE=1                      # -> fa2_triton_b200/libfa2_sm100_trace.so
    FA2_SM100_LIB=fa2_triton_b200/libfa2_sm100_trace.so python scripts/timeline_trace.py [spill|det|atomic]

The kernels of that build stamp `[role][iteration][event]` slots of a device array (`FA2_STAMP` in csrc/*.cu; CTA 0 of the
forward = the heaviest tile pair under causal masking); `fa2_sm100_debug_fetch_trace(which, host, n)` copies them out
(which = 0 forward, 1 fused backward, 2 recomputing dQ, 3 dK/dV). Cycles are printed relative to the first stamp.
`profiles/r01_trace_{spill,det,atomic}.txt` were made with the first version of this script (lost with its container:
its name matched a .gitignore pattern and it was never committed); the forward's event names are the same, the backward
kernels are printed by slot number.
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

FWD_EVENTS = {0: "wait", 1: "s_full", 2: "ld_done", 3: "max_done", 4: "exp_done", 5: "p_arrive"}
FWD_MMA_EVENTS = {0: "p0_seen", 1: "pv0_issued", 2: "k_ready0", 3: "qk0_issued",
                  4: "p1_seen", 5: "pv1_issued", 6: "k_ready1", 7: "qk1_issued"}
MODES = {"spill": True, "det": 1, "atomic": False}  # what set_deterministic() takes (backward.py)


def fetch(lib, which, roles):
    n = roles * 64 * 8
    buf = (ctypes.c_longlong * n)()
    fn = lib.fa2_sm100_debug_fetch_trace
    fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_size_t]
    fn.restype = ctypes.c_int
    rc = fn(which, buf, n)
    if rc != 0:
        raise RuntimeError(f"fa2_sm100_debug_fetch_trace({which}) returned {rc}")
    return [[[buf[(r * 64 + it) * 8 + ev] for ev in range(8)] for it in range(64)] for r in range(roles)]


def show(title, trace, role_names, event_names):
    stamps = [v for role in trace for it in role for v in it if v]
    if not stamps:
        print(f"==== {title}: no stamps (kernel not launched in this mode) ====")
        return
    t0 = min(stamps)
    print(f"==== {title} ====")
    for it in range(64):
        if not any(v for role in trace for v in role[it]):
            continue
        print(f"-- iteration {it}")
        for r, role in enumerate(trace):
            if not any(role[it]):
                continue
            names = event_names[r] if isinstance(event_names, list) else event_names
            cells = "  ".join(f"{names.get(ev, f'ev{ev}')}={role[it][ev] - t0}" for ev in range(8) if role[it][ev])
            print(f"   {role_names[r]:8s} {cells}")


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "spill"
    lib = _lib.load() if hasattr(_lib, "load") else ctypes.CDLL(os.environ["FA2_SM100_LIB"])
    if not hasattr(lib, "fa2_sm100_debug_fetch_trace"):
        sys.exit("this library was not built with TRACE=1 (no fa2_sm100_debug_fetch_trace)")
    dev = "cuda:0"
    B, H, S, D = 1, 2, 4096, 128
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = (torch.empty(B, S, H, D, device=dev).normal_(0, 0.5, generator=g).bfloat16().requires_grad_(True) for _ in range(3))
    do = torch.empty(B, S, H, D, device=dev).normal_(0, 1.0, generator=g).bfloat16()
    prev = fa2.is_deterministic()
    fa2.set_deterministic(MODES[mode])
    try:
        for _ in range(2):  # the second pass overwrites the stamps of the cold one
            out = fa2.flash_attn_func(q, k, v, causal=True)
            torch.autograd.grad(out, (q, k, v), do)
        torch.cuda.synchronize()
    finally:
        fa2.set_deterministic(prev)
    show("forward (first CTA = last tile pair)", fetch(lib, 0, 3), ["wg0", "wg1", "mma"], [FWD_EVENTS, FWD_EVENTS, FWD_MMA_EVENTS])
    for which, name in ((3, "dK/dV kernel"), (2, "recomputing dQ kernel"), (1, "fused backward kernel")):
        show(name, fetch(lib, which, 4), [f"role{r}" for r in range(4)], {})


if __name__ == "__main__":
    main()
