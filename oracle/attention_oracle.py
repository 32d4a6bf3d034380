"""numpy restatement of the reference's attention arithmetic. TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function cites the reference lines it follows (paths relative to remi-or/fa2_triton). Everything is computed
per (batch, head) slice with explicit formulas - no autograd - in float64 by default, so it is an independent
check of both the reference's eager PyTorch oracle (src/reference_implementation.py) and the CUDA kernels.

Pinning: the reference holds no golden vectors; this file is pinned by
``tests/golden/*.npz``, produced by running the reference's own ``flash_attn_reference`` + autograd in the build
container (``oracle/gen_golden.py``), and by ``tests/golden/triton_*.npz``: O, LSE, dQ, dK, dV of the reference's
own Triton kernels executed by Triton's CPU interpreter (``oracle/gen_golden_triton.py``; fp16, incl. attention_mask,
bias and forward dropout). Both are checked in ``tests/test_oracle.py``. The LSE has no runnable reference test
(tests/test_logsumexp.py is dead code); it is pinned by the Triton kernel's own LSE output in those fixtures.
"""
from __future__ import annotations

import math

import numpy as np

LOG2E = 1.4426950408889634


def visibility_mask(seqlen_q: int, seqlen_k: int, causal: bool) -> np.ndarray:
    """Boolean [Sq, Sk], True where key j is visible to query i.

    Follows ``construct_local_mask`` with window (-1, 0): masked iff ``col > row + sk - sq``
    (src/reference_implementation.py:8-29, 74-75) - bottom-right aligned, identical to the kernel-side test
    ``offs_m >= offs_n - seqlen_k + seqlen_q`` (src/forward/compute_row_blocks.py:64-66).
    """
    if not causal:
        return np.ones((seqlen_q, seqlen_k), dtype=bool)
    row = np.arange(seqlen_q)[:, None]
    col = np.arange(seqlen_k)[None, :]
    return col <= row + (seqlen_k - seqlen_q)


def _expand_kv(x: np.ndarray, nheads_q: int) -> np.ndarray:
    """GQA: repeat each kv head for its query-head group (src/reference_implementation.py:80-81)."""
    rep = nheads_q // x.shape[2]
    return np.repeat(x, rep, axis=2) if rep > 1 else x


def _masked_scores(q, kx, causal, scale, bias, seqlens):
    """Scaled scores [B,H,Sq,Sk] with hidden entries at -inf, plus the bool [B,1,Sq,1] of live query rows.

    Order of operations as src/reference_implementation.py:84-102: key padding (:91-92), causal / local mask
    (:93-102; with per-sequence lengths the diagonal follows them, :16-29 - query and key lengths coincide in the
    reference's attention_mask mode, src/forward/caller.py:27-28), then the additive bias (:101-102; the kernel adds
    it to the scaled scores too: src/forward/compute_row_blocks.py:68-69 multiplies by log2(e)/scale before the
    common scale factor)."""
    B, Sq, H, D = q.shape
    Sk = kx.shape[1]
    s = np.einsum("bthd,bshd->bhts", q, kx) * scale
    s = np.where(visibility_mask(Sq, Sk, causal)[None, None], s, -np.inf)
    live = np.ones((B, 1, Sq, 1), dtype=bool)
    if seqlens is not None:
        assert Sq == Sk, "attention_mask mode needs seqlen_q == seqlen_k (src/forward/caller.py:28)"
        n = np.asarray(seqlens).reshape(B, 1, 1, 1)
        s = np.where(np.arange(Sk).reshape(1, 1, 1, Sk) < n, s, -np.inf)
        live = np.arange(Sq).reshape(1, 1, Sq, 1) < n
    if bias is not None:
        s = s + np.asarray(bias, dtype=s.dtype)
    return s, live


def _softmax(s):
    m = s.max(axis=-1, keepdims=True)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(s - m_safe)
    l = e.sum(axis=-1, keepdims=True)
    p = np.where(l > 0, e / np.where(l > 0, l, 1.0), 0.0)
    return p, m_safe, l


def attention_forward(q, k, v, causal=False, softmax_scale=None, dtype=np.float64, bias=None, seqlens=None,
                      dropout_keep=None, dropout_p=0.0):
    """O and log2-domain LSE.

    Optional: ``bias`` broadcastable to [B,Hq,Sq,Sk] (added to the scaled scores), ``seqlens`` [B] (right-padded
    batch: keys beyond a sequence's length are hidden, its padded query rows give O = 0 and LSE = 0 like the
    zero-initialised buffers of src/forward/caller.py:70-74), ``dropout_keep`` bool [B,Hq,Sq,Sk] with ``dropout_p``
    (probabilities of dropped entries are zeroed AFTER the row sum and the result is scaled by 1/(1-p):
    src/forward/compute_row_blocks.py:73-79, src/forward/kernel.py:256-257; the LSE is that of the undropped row).

    q: [B, Sq, Hq, D]; k, v: [B, Sk, Hkv, D]. Returns (o [B, Sq, Hq, D], lse2 [B, Hq, Sq]) in ``dtype``.
    Scores/softmax/output follow src/reference_implementation.py:84-120 (scale 1/sqrt(d) default as
    src/forward/caller.py:42; rows with no visible key give 0, :107-108). LSE follows the kernel's definition
    ``m + log2(l)`` over scores multiplied by ``softmax_scale * log2(e)`` (src/forward/kernel.py:119,
    src/forward/compute_row_blocks.py:99-101; checked against tests/test_logsumexp.py:53-74's intent);
    fully masked rows get -inf.
    """
    q = np.asarray(q, dtype=dtype)
    k = _expand_kv(np.asarray(k, dtype=dtype), q.shape[2])
    v = _expand_kv(np.asarray(v, dtype=dtype), q.shape[2])
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    scale = 1.0 / math.sqrt(D) if softmax_scale is None else softmax_scale
    s, live = _masked_scores(q, k, causal, scale, bias, seqlens)
    p, m_safe, l = _softmax(s)
    with np.errstate(divide="ignore"):
        lse2 = (m_safe[..., 0] + np.log(l[..., 0])) * LOG2E
    lse2 = np.where(l[..., 0] > 0, lse2, -np.inf)
    lse2 = np.where(live[..., 0], lse2, 0.0)
    p = np.where(live, p, 0.0)
    if dropout_keep is not None:
        p = np.where(dropout_keep, p, 0.0) / (1.0 - dropout_p)
    o = np.einsum("bhts,bshd->bthd", p, v)
    return o.astype(dtype), lse2.astype(dtype)


def attention_backward(q, k, v, do, causal=False, softmax_scale=None, dtype=np.float64, bias=None, seqlens=None):
    """dQ, dK, dV by the closed-form softmax-attention gradient (``bias`` / ``seqlens`` as in the forward; the
    bias gets no gradient, like the reference: src/wrapper.py:86).

    Same maths the reference kernels implement: P from the LSE, delta = rowsum(O*dO)
    (src/backward/compute_delta.py:59-71), dV = P^T dO, dP = dO V^T, dS = P*(dP - delta)*scale,
    dK = dS^T Q, dQ = dS K (src/backward/compute_dkdv.py:71-110, src/backward/compute_dq.py:53-76) and the GQA
    head-group sum (src/backward/caller.py:163-165). It equals autograd through
    src/reference_implementation.py:84-120, which is what tests/golden pins.
    Returns (dq [B,Sq,Hq,D], dk [B,Sk,Hkv,D], dv [B,Sk,Hkv,D]).
    """
    q = np.asarray(q, dtype=dtype)
    k0 = np.asarray(k, dtype=dtype)
    v0 = np.asarray(v, dtype=dtype)
    do = np.asarray(do, dtype=dtype)
    B, Sq, H, D = q.shape
    Sk, Hkv = k0.shape[1], k0.shape[2]
    kx, vx = _expand_kv(k0, H), _expand_kv(v0, H)
    scale = 1.0 / math.sqrt(D) if softmax_scale is None else softmax_scale
    s, live = _masked_scores(q, kx, causal, scale, bias, seqlens)
    p, _, _ = _softmax(s)  # [B,H,Sq,Sk]
    p = np.where(live, p, 0.0)
    do = np.where(live.transpose(0, 2, 1, 3), do, 0.0)  # gradients arriving at padded rows are ignored
    o = np.einsum("bhts,bshd->bthd", p, vx)
    delta = np.einsum("bthd,bthd->bht", o, do)  # [B,H,Sq]
    dv = np.einsum("bhts,bthd->bshd", p, do)
    dp = np.einsum("bthd,bshd->bhts", do, vx)
    ds = p * (dp - delta[..., None]) * scale
    dq = np.einsum("bhts,bshd->bthd", ds, kx)
    dk = np.einsum("bhts,bthd->bshd", ds, q)
    if Hkv != H:
        g = H // Hkv
        dk = dk.reshape(B, Sk, Hkv, g, D).sum(axis=3)
        dv = dv.reshape(B, Sk, Hkv, g, D).sum(axis=3)
    return dq.astype(dtype), dk.astype(dtype), dv.astype(dtype)


def generate_test_data(batch, nheads_q, nheads_kv, seqlen_q, seqlen_k, head_dim, dtype="float32", seed=0):
    """Seeded inputs with the reference's distributions: Q, K, V ~ N(0, 0.5^2), dO ~ N(0, 1)
    (tests/utils.py:21-25). Uses numpy's Generator so the same arrays can be rebuilt without torch."""
    rng = np.random.default_rng(seed)
    q = rng.normal(0.0, 0.5, size=(batch, seqlen_q, nheads_q, head_dim)).astype(dtype)
    k = rng.normal(0.0, 0.5, size=(batch, seqlen_k, nheads_kv, head_dim)).astype(dtype)
    v = rng.normal(0.0, 0.5, size=(batch, seqlen_k, nheads_kv, head_dim)).astype(dtype)
    do = rng.normal(0.0, 1.0, size=(batch, seqlen_q, nheads_q, head_dim)).astype(dtype)
    return q, k, v, do
