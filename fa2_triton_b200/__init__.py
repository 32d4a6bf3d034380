"""fa2_triton_b200 - FlashAttention-2 forward/backward for NVIDIA B200 (sm_100a) behind the
``flash_attn_func`` / ``FlashAttnFunc`` entry point of remi-or/fa2_triton (reference src/__init__.py:1-4).

The reference also re-exports its eager-PyTorch oracle ``flash_attn_reference``; here the oracle is test
infrastructure and lives in ``oracle/`` at the repository root, never on the product path.
"""
from .backward import _flash_attn_backward, is_deterministic, set_deterministic
from .forward import _flash_attn_forward
from .wrapper import FlashAttnFunc, flash_attn_func

__all__ = [
    "flash_attn_func",
    "FlashAttnFunc",
    "_flash_attn_forward",
    "_flash_attn_backward",
    "set_deterministic",
    "is_deterministic",
]
