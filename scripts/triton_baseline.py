#!/usr/bin/env python
"""The reference's own Triton kernels on the same B200, next to this repo's CUDA path.

    python scripts/triton_baseline.py [--out profiles/r02_triton_baseline.json] [--skip-tests]

Needs `baseline/_ref` = an unmodified copy of the reference tree (git-ignored; it travels to the GPU box with the
gpurun snapshot). Nothing of it is imported by the product; this script is measurement / parity tooling.

(a) timing: forward (BASELINE config 2) and forward + backward (config 3) of `src.wrapper.flash_attn_func`
    (unmodified, imported from baseline/_ref after the two import shims of SURVEY F2 / F4), autotuner warmed first,
    CUDA events around back-to-back steps, same inputs / seed / warm-up as bench.py; the same loop for this repo's entry
    point in the same process, same box.
(b) parity in bf16 on identical tensors: O, LSE, dQ, dK, dV of both implementations at config 3 and the ragged S = 1000
    cases of config 4, against the north-star tolerance (2e-2 abs + 1e-2 rel; LSE 1e-3), and both against the fp32
    eager oracle on sampled (b, h) slices.
(c) with --tests: the reference's OWN pytest files (tests/test_fwd_bwd.py, tests/test_fwd_only.py) with
    `flash_attn_func` replaced by this repo's entry point (see scripts/run_reference_tests.py).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FA2_REFERENCE_DIR", ROOT / "baseline" / "_ref"))
sys.path.insert(0, str(ROOT))


def import_reference():
    """SURVEY F2 + F4 shims, then `import src` from the unmodified tree."""
    import triton.language as tl

    if not hasattr(tl, "const_pointer_type"):
        tl.const_pointer_type = tl.pointer_type
    t = types.ModuleType("tests")
    t.__path__ = [str(REF / "tests")]
    sys.modules["tests"] = t
    sys.path.insert(0, str(REF))
    import src  # noqa: F401
    from src.backward.caller import _flash_attn_backward
    from src.forward.caller import _flash_attn_forward
    from src.wrapper import flash_attn_func

    return flash_attn_func, _flash_attn_forward, _flash_attn_backward


def fa_flops(B, H, Sq, Sk, D, causal, bwd):
    f = 4.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
    return f * (3.5 if bwd else 1.0)


def make_inputs(torch, B, H, S, D, dev, seed=0, dtype=None):
    dtype = dtype or torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(seed)
    out = []
    for std in (0.5, 0.5, 0.5, 1.0):
        t = torch.empty((B, S, H, D), dtype=dtype, device=dev)
        for b in range(B):
            t[b] = torch.empty((S, H, D), dtype=torch.float32, device=dev).normal_(0, std, generator=g).to(dtype)
        out.append(t)
    return out


def time_steps(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def autotune_winners():
    out = {}
    try:
        from src.backward.compute_delta import _compute_delta
        from src.backward.kernel import _bwd_kernel
        from src.forward.kernel import _fwd_kernel

        for name, k in (("_fwd_kernel", _fwd_kernel), ("_compute_delta", _compute_delta), ("_bwd_kernel", _bwd_kernel)):
            cache = getattr(k, "cache", None)
            if cache:
                out[name] = {str(key): str(cfg) for key, cfg in cache.items()}
    except Exception as e:  # the autotuner's internals are not an API
        out["error"] = repr(e)
    return out


def err_table(torch, got, want, atol=2e-2, rtol=1e-2):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    return {"max_abs": err.max().item(), "mean_abs": err.mean().item(),
            "worst_excess_over_tol": (err - bound).max().item(), "within_tol": bool((err <= bound).all().item())}


def parity_case(torch, ours_mod, ref_fwd, ref_bwd, B, H, S, D, causal, dev, oracle_slices=2):
    """Both implementations through their two-function boundary on identical bf16 tensors."""
    from oracle import eager_reference as er

    q, k, v, do = make_inputs(torch, B, H, S, D, dev)
    res = {"shape": f"B={B},H={H},S={S},D={D},causal={causal},bf16"}
    o_r, lse_r, scale, seed = ref_fwd(q, k, v, None, None, 0.0, causal, None, None)
    dq_r, dk_r, dv_r = ref_bwd(do, q, k, v, None, None, o_r, lse_r, 0.0, causal, scale, seed)
    o_o, lse_o, scale_o, seed_o = ours_mod._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    dq_o, dk_o, dv_o = ours_mod._flash_attn_backward(do, q, k, v, None, None, o_o, lse_o, 0.0, causal, scale_o, seed_o)
    torch.cuda.synchronize()
    res["ours_vs_triton"] = {
        "O": err_table(torch, o_o, o_r), "dQ": err_table(torch, dq_o.to(q.dtype), dq_r.to(q.dtype)),
        "dK": err_table(torch, dk_o, dk_r), "dV": err_table(torch, dv_o, dv_r),
        "LSE": err_table(torch, lse_o[:, :, :S], lse_r[:, :, :S], atol=1e-3, rtol=0.0),
    }
    # both against the fp32 eager oracle on sampled (b, h) slices (the oracle materialises S x S scores)
    vs_o = {"ours": {}, "triton": {}}
    picks = [(0, 0), (B - 1, H - 1)][:oracle_slices]
    for (b, h) in picks:
        sl = lambda t: t[b:b + 1, :, h:h + 1].float()  # noqa: E731
        qq, kk, vv = (sl(t).clone().requires_grad_(True) for t in (q, k, v))
        o_ref = er.flash_attn_reference(qq, kk, vv, causal=causal)
        g = torch.autograd.grad(o_ref, (qq, kk, vv), sl(do))
        for name, mine, theirs, want in (("O", o_o, o_r, o_ref), ("dQ", dq_o, dq_r, g[0]), ("dK", dk_o, dk_r, g[1]),
                                         ("dV", dv_o, dv_r, g[2])):
            for who, t in (("ours", mine), ("triton", theirs)):
                e = err_table(torch, sl(t.to(q.dtype)), want)
                cur = vs_o[who].get(name)
                if cur is None or e["max_abs"] > cur["max_abs"]:
                    vs_o[who][name] = e
    res["vs_fp32_oracle_sampled"] = vs_o
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "r02_triton_baseline.json"))
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    result = {"reference_dir": str(REF), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    if not (REF / "src" / "wrapper.py").exists():
        result["unavailable"] = f"{REF} does not hold the reference tree"
        print(json.dumps(result))
        return 0
    import torch
    import triton

    result["triton"] = triton.__version__
    result["torch"] = torch.__version__
    result["gpu"] = torch.cuda.get_device_name(0)
    dev = torch.device("cuda", 0)
    os.environ.setdefault("TRITON_PRINT_AUTOTUNING", "1")
    try:
        ref_func, ref_fwd, ref_bwd = import_reference()
    except Exception as e:
        result["unavailable"] = "import of the reference failed: " + repr(e)
        print(json.dumps(result))
        return 0
    import fa2_triton_b200 as fa2
    from fa2_triton_b200 import backward as ours_b, forward as ours_f

    ours_mod = types.SimpleNamespace(_flash_attn_forward=ours_f._flash_attn_forward,
                                     _flash_attn_backward=ours_b._flash_attn_backward)
    fa2.set_deterministic(True)

    # ------------------------------------------------ (a) timing --------------------------------------------------
    timing = {}
    for name, (B, H, S, D, causal, bwd) in {"cfg2_fwd": (4, 32, 4096, 128, True, False),
                                            "cfg3_fwd_bwd": (4, 32, 4096, 128, True, True),
                                            "cfg4_fwd_bwd": (2, 16, 16384, 64, False, True)}.items():
        q, k, v, do = make_inputs(torch, B, H, S, D, dev)
        fl = fa_flops(B, H, S, S, D, causal, bwd)

        def step(func):
            if bwd:
                qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
                out = func(qq, kk, vv, causal=causal)
                return torch.autograd.grad(out, (qq, kk, vv), do)
            return func(q, k, v, causal=causal)

        row = {}
        for who, func in (("triton_reference", ref_func), ("ours", fa2.flash_attn_func)):
            try:
                t0 = time.time()
                step(func)  # first call: autotune (reference) / lazy init
                torch.cuda.synchronize()
                first = time.time() - t0
                ms = time_steps(torch, lambda: step(func), args.steps, args.warmup)
                row[who] = {"ms_per_step": ms, "tflops": fl / (ms * 1e-3) / 1e12, "first_call_s": first}
            except Exception as e:
                row[who] = {"error": repr(e)[:2000]}
        if "ms_per_step" in row.get("triton_reference", {}) and "ms_per_step" in row.get("ours", {}):
            row["speedup_ours_over_triton"] = row["triton_reference"]["ms_per_step"] / row["ours"]["ms_per_step"]
        timing[name] = row
        print(name, json.dumps(row), flush=True)
        del q, k, v, do
        torch.cuda.empty_cache()
    result["timing"] = timing
    result["timing_method"] = (f"CUDA events around {args.steps} back-to-back steps after {args.warmup} warm-up steps "
                               "(+ the autotuning first call), inputs resident in HBM, same process and box for both")
    result["autotune_winners"] = autotune_winners()

    # ------------------------------------------------ (b) parity --------------------------------------------------
    if not args.no_parity:
        parity = []
        for (B, H, S, D, causal) in [(4, 32, 4096, 128, True), (2, 16, 1000, 64, False), (2, 16, 1000, 64, True),
                                     (2, 4, 1000, 128, True)]:
            try:
                parity.append(parity_case(torch, ours_mod, ref_fwd, ref_bwd, B, H, S, D, causal, dev))
            except Exception as e:
                parity.append({"shape": f"B={B},H={H},S={S},D={D},causal={causal}", "error": repr(e)[:2000]})
            print(json.dumps(parity[-1]), flush=True)
            torch.cuda.empty_cache()
        result["parity_bf16"] = parity
        result["parity_tolerance"] = "north star: |x - ref| <= 2e-2 + 1e-2 |ref| for O, dQ, dK, dV (bf16); LSE abs 1e-3"
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(result, indent=1))
    print("wrote", args.out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
