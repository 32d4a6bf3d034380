"""Generate tests/golden/triton_*.npz by running the REFERENCE'S OWN TRITON KERNELS on the CPU.

There is no GPU in the build container and `/root/reference` does not exist on the GPU box, but Triton ships an
interpreter (`TRITON_INTERPRET=1`) that executes `@triton.jit` functions with numpy. This script imports the reference
unmodified (`src.forward.caller._flash_attn_forward`, `src.backward.caller._flash_attn_backward`: the two functions
behind `src/wrapper.py:13-86`) and drives them on CPU tensors. Three things are needed for that, none of which changes
what the kernels compute:

* `tl.const_pointer_type` does not exist in Triton 3.6: aliased to `tl.pointer_type` before the import.
* The callers assert `q.is_cuda` (`src/forward/caller.py:41`, `src/backward/caller.py:53`): the inputs are CPU tensors of
  a `torch.Tensor` subclass whose `is_cuda` answers True.
* `@triton.autotune` would benchmark its candidate configurations on a GPU: each autotuner's candidate list is
  narrowed (in memory) to the one configuration of `src/forward/kernel.py:36-41` / `src/backward/kernel.py:36-40` /
  `src/backward/compute_delta.py:9-14` that the reference's own pruning rule keeps for the shape and that a B200 can
  launch (128-wide tiles when the sequence allows it).

The interpreter has no bfloat16 (numpy lacks the type; its results are garbage), so every case is fp16 - the
kernels' arithmetic is the same for both 16-bit types apart from the final roundings. What is stored per case:
inputs (q, k, v, do and, where used, attention_mask / bias / dropout seed) and the Triton kernel's O, LSE (log2
domain, `[:, :, :Sq]`), dQ (fp32, as `src/backward/caller.py:86` produces it), dK, dV. The dropout case also stores the
keep-mask of the reference's test helper (`tests/utils.py:169-207`, `tl.rand(seed, linear offset) > p`), which pins the
Philox stream the forward kernel draws from (`src/forward/compute_row_blocks.py:76-79`).

Run in the build container (reference at /root/reference or $FA2_REFERENCE_DIR):
    python oracle/gen_golden_triton.py
The .npz files are committed; nothing at test time needs the reference checkout or Triton.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

os.environ["TRITON_INTERPRET"] = "1"  # must be set before triton is imported

import numpy as np  # noqa: E402
import torch  # noqa: E402
import triton  # noqa: E402
import triton.language as tl  # noqa: E402

REF = Path(os.environ.get("FA2_REFERENCE_DIR", "/root/reference"))
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


class _CudaLabelled(torch.Tensor):
    """CPU tensor that passes the reference's `assert q.is_cuda`."""
    is_cuda = property(lambda self: True)


def _import_reference():
    if not hasattr(tl, "const_pointer_type"):
        tl.const_pointer_type = tl.pointer_type
    sys.path.insert(0, str(REF))
    # a site-packages `tests` package shadows the reference's tests/ directory
    pkg = types.ModuleType("tests")
    pkg.__path__ = [str(REF / "tests")]
    sys.modules["tests"] = pkg
    from src.backward import compute_delta as cd
    from src.backward.caller import _flash_attn_backward
    from src.backward.kernel import _bwd_kernel
    from src.forward.caller import _flash_attn_forward
    from src.forward.kernel import _fwd_kernel
    from tests.utils import _generate_dropout_mask
    return _flash_attn_forward, _flash_attn_backward, _fwd_kernel, _bwd_kernel, cd._compute_delta, _generate_dropout_mask


class _Pinned:
    """Narrow an Autotuner's candidates to one configuration for the duration of a call."""

    def __init__(self, kern):
        self.kern, self.all = kern, list(kern.configs)

    def pin(self, pred):
        keep = [c for c in self.all if pred(c.kwargs)]
        assert len(keep) == 1, keep
        self.kern.configs = keep


def _tile(n):  # the largest of the reference's square tile sizes that its pruning rule keeps and sm_100a can launch
    return 128 if n >= 128 else 64 if n >= 64 else 32 if n >= 32 else 16


# name: B, Sq, Sk, Hq, Hkv, D, causal, extras
CASES = {
    "triton_causal_fp16_b1h2s128d64": (1, 128, 128, 2, 2, 64, True, {}),
    "triton_causal_fp16_b1h2s256d128": (1, 256, 256, 2, 2, 128, True, {}),
    "triton_ragged_gqa_fp16_b2h4kv2sq113sk203d40": (2, 113, 203, 4, 2, 40, False, {}),
    "triton_causal_sq_lt_sk_fp16_b1h2sq65sk200d128": (1, 65, 200, 2, 1, 128, True, {}),
    # several 128-wide key blocks per query block with the B200-sized tiles: the online-softmax rescale path
    "triton_causal_multiblock_gqa_fp16_b1h2kv1sq384sk640d128": (1, 384, 640, 2, 1, 128, True, {}),
    "triton_causal_sq_gt_sk_fp16_b1h2sq200sk72d64": (1, 200, 72, 2, 2, 64, True, {}),
    "triton_varlen_causal_fp16_b3h2s96d64": (3, 96, 96, 2, 2, 64, True, {"lengths": (96, 70, 1)}),
    "triton_bias_fp16_b2h2sq108sk256d64": (2, 108, 256, 2, 2, 64, False, {"bias": True}),
    "triton_bias_causal_fp16_b1h3sq127sk130d32": (1, 127, 130, 3, 3, 32, True, {"bias": True}),
    "triton_dropout_fwd_fp16_b2h3sq128sk217d64": (2, 128, 217, 3, 3, 64, False, {"dropout_p": 0.1, "seed": 20261018}),
    "triton_dropout_causal_fwd_fp16_b1h2sq100sk100d64": (1, 100, 100, 2, 2, 64, True, {"dropout_p": 0.17, "seed": 7}),
    # dropout together with attention_mask: the counters are numbered over the PACKED batch (src/forward/kernel.py:105-112,
    # 147-149). The reference's test helper cannot build this mask (it numbers [B, H, Sq, Sk]), so only O and the LSE are kept.
    "triton_varlen_dropout_fwd_fp16_b3h2s96d64": (3, 96, 96, 2, 2, 64, False,
                                                   {"lengths": (96, 57, 20), "dropout_p": 0.2, "seed": 424242}),
    "triton_varlen_dropout_causal_fwd_fp16_b3h2s80d64": (3, 80, 80, 2, 2, 64, True,
                                                          {"lengths": (33, 80, 64), "dropout_p": 0.1, "seed": 99}),
}


def main():
    if not REF.exists():
        sys.exit(f"reference checkout not found at {REF}")
    fwd, bwd, k_fwd, k_bwd, k_delta, k_mask = _import_reference()
    p_fwd, p_bwd, p_delta = _Pinned(k_fwd), _Pinned(k_bwd), _Pinned(k_delta)
    OUT.mkdir(parents=True, exist_ok=True)
    dt = torch.float16
    for name, (B, Sq, Sk, Hq, Hkv, D, causal, extra) in CASES.items():
        if (OUT / f"{name}.npz").exists() and "--all" not in sys.argv:
            print(f"{name}: exists (pass --all to regenerate)")
            continue
        # same recipe as the reference's tests/utils.py:21-25 (seed 0; Q,K,V ~ N(0, 0.5), dO ~ N(0, 1))
        torch.manual_seed(0)
        q = torch.empty(B, Sq, Hq, D, dtype=dt).normal_(0, 0.5)
        k = torch.empty(B, Sk, Hkv, D, dtype=dt).normal_(0, 0.5)
        v = torch.empty(B, Sk, Hkv, D, dtype=dt).normal_(0, 0.5)
        do = torch.randn(B, Sq, Hq, D, dtype=dt)
        mask = bias = None
        if "lengths" in extra:  # right-padded batch, like tests/utils.py generate_attention_mask
            mask = torch.zeros(B, Sq, dtype=torch.bool)
            for b, n in enumerate(extra["lengths"]):
                mask[b, :n] = True
        if extra.get("bias"):  # tests/core.py:28
            bias = torch.rand(1, 1, Sq, Sk, dtype=dt)
        dropout_p, seed = extra.get("dropout_p", 0.0), extra.get("seed")

        t = _tile(min(Sq, Sk)) if mask is None else _tile(min(extra["lengths"]) if min(extra["lengths"]) >= 16 else 64)
        t = min(t, _tile(min(Sq, Sk)))
        p_fwd.pin(lambda kw: kw["BLOCK_M"] == max(t, 32))
        bwd_cfg = {128: (64, 128, 128, 64), 64: (64, 64, 64, 64), 32: (32, 16, 16, 32), 16: (16, 16, 16, 16)}[t]
        p_bwd.pin(lambda kw: (kw["BLOCK_M1"], kw["BLOCK_N1"], kw["BLOCK_M2"], kw["BLOCK_N2"]) == bwd_cfg)
        p_delta.pin(lambda kw: kw["BLOCK_M"] == max(t, 16))

        lab = lambda x: None if x is None else x.as_subclass(_CudaLabelled)
        o, lse, scale, seed_used = fwd(lab(q), lab(k), lab(v), lab(mask), lab(bias), dropout_p, causal, None, seed)
        saved = dict(q=q.numpy(), k=k.numpy(), v=v.numpy(), do=do.numpy(), out=torch.Tensor(o).numpy(),
                     lse=torch.Tensor(lse)[:, :, :Sq].numpy(), causal=np.array(causal),
                     softmax_scale=np.array(scale, dtype=np.float64), tile=np.array(t))
        if mask is not None:
            saved["attention_mask"] = mask.numpy()
        if bias is not None:
            saved["bias"] = bias.numpy()
        if dropout_p > 0 and mask is not None:
            saved.update(dropout_p=np.array(dropout_p), dropout_seed=np.array(seed_used, dtype=np.int64))
        elif dropout_p > 0:
            # the keep-mask exactly as the reference's tests build it (tests/utils.py:169-207)
            keep = torch.zeros(B, Hq, Sq, Sk)
            n = keep.numel()
            k_mask[(triton.cdiv(n, 64),)](keep, seed_used, dropout_p, n, 64)
            saved.update(dropout_p=np.array(dropout_p), dropout_seed=np.array(seed_used, dtype=np.int64),
                         dropout_keep=np.packbits(keep.bool().numpy().reshape(-1)))
        else:  # the reference's backward raises NotImplementedError under dropout (src/utils.py:87-88)
            dq, dk, dv = bwd(lab(do), lab(q), lab(k), lab(v), lab(bias), lab(mask), o, lse, 0.0, causal, scale, seed_used)
            assert dq.dtype == torch.float32
            saved.update(dq=torch.Tensor(dq).numpy(), dk=torch.Tensor(dk).numpy(), dv=torch.Tensor(dv).numpy())
        np.savez_compressed(OUT / f"{name}.npz", **saved)
        print(f"{name}: tile {t}, sum(out)={torch.Tensor(o).double().sum().item():.5f} "
              f"sum(lse)={torch.Tensor(lse)[:, :, :Sq].double().nan_to_num(0, 0, 0).sum().item():.5f}"
              + ("" if dropout_p > 0 else f" sum(dq)={torch.Tensor(dq).double().sum().item():.5f}"))


if __name__ == "__main__":
    main()
