"""numpy restatement of the random stream behind the reference's forward dropout. TEST INFRASTRUCTURE ONLY.

The reference draws ``tl.rand(dropout_seed, offset) > dropout_p`` per score (src/forward/compute_row_blocks.py:76-79)
with ``offset`` = the element's linear index in a ``[B, Hq, Sq, Sk]`` tensor (src/forward/kernel.py:147-149; its tests
build the same mask with tests/utils.py:169-207). ``tl.rand`` is Triton's (triton==3.6.0, ``triton/language/random.py``,
third-party and not vendored by the reference): Philox-4x32 with 10 rounds, counter ``(offset, 0, 0, 0)``, key = the
64-bit seed split into (low, high) words; the first output word, reinterpreted as int32 ``x``, becomes the uniform
``(x if x >= 0 else -x - 1) * 4.6566127342e-10`` in fp32. Pinned by the keep-masks stored in
``tests/golden/triton_dropout_*.npz`` (produced by the reference's own helper under Triton's interpreter).
"""
from __future__ import annotations

import numpy as np

_ROUND_A, _ROUND_B = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_KEY_A, _KEY_B = 0x9E3779B9, 0xBB67AE85
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10_first_word(seed: int, offsets: np.ndarray) -> np.ndarray:
    """First output word (uint32) of Philox-4x32-10 for counters (offset, 0, 0, 0) and key (seed_lo, seed_hi)."""
    c0 = np.asarray(offsets).astype(np.uint64) & _M32
    c1 = np.zeros_like(c0)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        pa, pb = _ROUND_A * c0, _ROUND_B * c2  # 32 x 32 -> 64 bit products
        c0, c1, c2, c3 = ((pb >> np.uint64(32)) ^ c1 ^ np.uint64(k0), pb & _M32,
                          (pa >> np.uint64(32)) ^ c3 ^ np.uint64(k1), pa & _M32)
        k0, k1 = (k0 + _KEY_A) & 0xFFFFFFFF, (k1 + _KEY_B) & 0xFFFFFFFF
    return c0.astype(np.uint32)


def uniform(seed: int, offsets: np.ndarray) -> np.ndarray:
    """``tl.rand(seed, offsets)``: fp32 in [0, 1)."""
    x = philox4x32_10_first_word(seed, offsets).view(np.int32)
    x = np.where(x < 0, -x - 1, x)
    return x.astype(np.float32) * np.float32(4.6566127342e-10)


def dropout_keep_mask(seed: int, dropout_p: float, B: int, H: int, Sq: int, Sk: int) -> np.ndarray:
    """bool [B, H, Sq, Sk], True where the probability is kept."""
    offs = np.arange(B * H * Sq * Sk, dtype=np.uint64).reshape(B, H, Sq, Sk)
    return uniform(seed, offs) > np.float32(dropout_p)


def dropout_keep_mask_varlen(seed: int, dropout_p: float, lengths, H: int, S: int) -> np.ndarray:
    """bool [B, H, S, S] keep-mask of a right-padded batch given through `attention_mask` (the reference's VARLEN mode):
    sequence b of length L that starts at row c of the PACKED batch draws tl.rand at offset
    L * (c + L * head) + row * L + col  (src/forward/kernel.py:105-112: `off_batch = 0`, `cu_seq_start_q`;
    :147-149: `dropout_off`), int32 arithmetic, i.e. modulo 2**32. Entries outside a sequence are True (unused)."""
    keep = np.ones((len(lengths), H, S, S), dtype=bool)
    start = 0
    for b, L in enumerate(int(n) for n in lengths):
        if L > 0:
            head = np.arange(H, dtype=np.uint64)[:, None, None]
            row = np.arange(L, dtype=np.uint64)[None, :, None]
            col = np.arange(L, dtype=np.uint64)[None, None, :]
            offs = (np.uint64(L) * (np.uint64(start) + np.uint64(L) * head) + row * np.uint64(L) + col) & np.uint64(0xFFFFFFFF)
            keep[b, :, :L, :L] = uniform(seed, offs) > np.float32(dropout_p)
        start += L
    return keep
