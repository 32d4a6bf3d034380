"""CPU oracle for the FlashAttention-2 hot path. TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of ``bench.py``
may import this package; nothing under ``fa2_triton_b200/`` does.
"""
