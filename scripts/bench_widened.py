"""Throughput of the rows SURVEY.md section 8(f) marks "next" - GQA, Sq != Sk, fp16, head dims below 128,
attention_mask, additive bias, forward dropout - measured like bench.py's `value`: whole fwd (+ bwd) steps through the
public API (`flash_attn_func` + `torch.autograd.grad`), inputs resident in HBM, CUDA events around `steps` back-to-back
steps after warm-up. FLOPs are the algorithmic ones of the VISIBLE score entries (4 * D * visible pairs per head, x 3.5
with the backward), so masked / padded work does not count.
Usage: python scripts/bench_widened.py [steps=20]  ->  one JSON line per case"""
import json
import math
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import fa2_triton_b200 as fa2

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda:0"


def visible_pairs(sq, sk, causal):
    if not causal:
        return sq * sk
    return sum(max(0, min(sk, i + sk - sq + 1)) for i in range(sq))


CASES = [
    # name, B, Sq, Sk, Hq, Hkv, D, dtype, causal, bwd, extras
    ("cfg3 baseline (MHA, bf16, causal)", 4, 4096, 4096, 32, 32, 128, torch.bfloat16, True, True, {}),
    ("GQA 32:8 heads", 4, 4096, 4096, 32, 8, 128, torch.bfloat16, True, True, {}),
    ("MQA 32:1 heads", 4, 4096, 4096, 32, 1, 128, torch.bfloat16, True, True, {}),
    ("fp16", 4, 4096, 4096, 32, 32, 128, torch.float16, True, True, {}),
    ("Sq=1024 < Sk=8192 causal (bottom-right)", 4, 1024, 8192, 32, 32, 128, torch.bfloat16, True, True, {}),
    ("ragged S=4001, head dim 96 (D=128 kernel, TMA zero-fill)", 4, 4001, 4001, 32, 32, 96, torch.bfloat16, True, True, {}),
    ("head dim 64, non-causal", 4, 4096, 4096, 32, 32, 64, torch.bfloat16, False, True, {}),
    ("attention_mask: lengths 4096 / 3000 / 2048 / 1000", 4, 4096, 4096, 32, 32, 128, torch.bfloat16, True, True,
     {"lengths": (4096, 3000, 2048, 1000)}),
    ("bias [1, 1, Sq, Sk] bf16", 4, 4096, 4096, 32, 32, 128, torch.bfloat16, True, True, {"bias": (1, 1)}),
    ("bias [B, H, Sq, Sk] bf16", 2, 2048, 2048, 16, 16, 128, torch.bfloat16, False, True, {"bias": (2, 16)}),
    ("dropout p=0.1, forward only", 4, 4096, 4096, 32, 32, 128, torch.bfloat16, True, False, {"dropout_p": 0.1}),
    ("forward only (cfg2)", 4, 4096, 4096, 32, 32, 128, torch.bfloat16, True, False, {}),
    ("head dim 256, causal", 4, 4096, 4096, 16, 16, 256, torch.bfloat16, True, True, {}),
    ("head dim 207 (zero-padded to 208), causal", 4, 4096, 4096, 16, 16, 207, torch.bfloat16, True, True, {}),
    ("head dim 256, forward only", 4, 4096, 4096, 16, 16, 256, torch.bfloat16, True, False, {}),
    ("cfg4: S=16384, head dim 64, non-causal", 2, 16384, 16384, 16, 16, 64, torch.bfloat16, False, True, {}),
    ("cfg4 ragged tail: S=1000, head dim 64, non-causal", 2, 1000, 1000, 16, 16, 64, torch.bfloat16, False, True, {}),
]

for name, B, Sq, Sk, Hq, Hkv, D, dtype, causal, bwd, extra in CASES:
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.empty(B, Sq, Hq, D, device=dev).normal_(0, 0.5, generator=g).to(dtype)
    k = torch.empty(B, Sk, Hkv, D, device=dev).normal_(0, 0.5, generator=g).to(dtype)
    v = torch.empty(B, Sk, Hkv, D, device=dev).normal_(0, 0.5, generator=g).to(dtype)
    do = torch.empty(B, Sq, Hq, D, device=dev).normal_(0, 1.0, generator=g).to(dtype)
    kw = dict(causal=causal)
    pairs = B * visible_pairs(Sq, Sk, causal)
    if "lengths" in extra:
        mask = torch.zeros(B, Sq, dtype=torch.bool, device=dev)
        for b, n in enumerate(extra["lengths"]):
            mask[b, :n] = True
        kw["attention_mask"] = mask
        pairs = sum(visible_pairs(n, n, causal) for n in extra["lengths"])
    if "bias" in extra:
        kw["attention_bias"] = torch.rand(extra["bias"][0], extra["bias"][1], Sq, Sk, device=dev, generator=g).to(dtype)
    if "dropout_p" in extra:
        kw.update(dropout_p=extra["dropout_p"], dropout_seed=1234)
    flops = 4.0 * D * pairs * Hq * (3.5 if bwd else 1.0)

    def step():
        if bwd:
            qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
            return torch.autograd.grad(fa2.flash_attn_func(qq, kk, vv, **kw), (qq, kk, vv), do)
        return fa2.flash_attn_func(q, k, v, **kw)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"case": name, "shape": f"B={B} Sq={Sq} Sk={Sk} Hq={Hq} Hkv={Hkv} D={D} {str(dtype)[6:]}"
                      f" causal={causal} {'fwd+bwd' if bwd else 'fwd'}", "ms_per_step": round(ms, 4),
                      "tflops": round(flops / ms / 1e9, 1), "steps": steps}), flush=True)
    del q, k, v, do, kw
    torch.cuda.empty_cache()
