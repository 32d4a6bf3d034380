"""Order in which the causal grids are worked through (csrc/fa2_kernels.cuh: tail_balanced_item - head-major, the last
heads block-major so that the last wave is made of the lightest blocks). Whatever the tail length, the order must be a
permutation of the grid: every (block, head, batch) exactly once. Host-side check through the test-only probe library;
no GPU needed."""
import ctypes
from pathlib import Path

import pytest


def probe_lib():
    path = Path(__file__).resolve().parent / "libfa2_probe.so"
    if not path.exists():
        pytest.skip("tests/libfa2_probe.so not built (make -C fa2_triton_b200/csrc)")
    lib = ctypes.CDLL(str(path))
    lib.fa2_probe_tail_item.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)]
    lib.fa2_probe_tail_item.restype = None
    lib.fa2_probe_tail_block.argtypes = [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_int)]
    lib.fa2_probe_tail_block.restype = None
    lib.fa2_probe_tail_log2_for.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.fa2_probe_tail_log2_for.restype = ctypes.c_int
    lib.fa2_probe_tail_heads_for.argtypes = [ctypes.c_longlong, ctypes.c_int]
    lib.fa2_probe_tail_heads_for.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("nx,ny,nb", [(32, 32, 4), (16, 32, 4), (1, 1, 1), (8, 3, 5), (64, 32, 32), (5, 7, 1), (128, 16, 2)])
@pytest.mark.parametrize("tail", [0, 1, 2, 3, 12, 24, 32, 1000])
def test_item_order_is_a_permutation_of_the_grid(nx, ny, nb, tail):
    lib = probe_lib()
    out = (ctypes.c_int * 3)()
    seen = set()
    order = []
    for lin in range(nx * ny * nb):
        lib.fa2_probe_tail_item(lin, nx, ny, ny * nb, tail, out)
        x, head, batch = out[0], out[1], out[2]
        assert 0 <= x < nx and 0 <= head < ny and 0 <= batch < nb
        seen.add((x, head, batch))
        order.append((x, head, batch))
    assert len(seen) == nx * ny * nb
    k = min(tail, ny * nb)
    if k < 2:
        # plain linear order: x fastest, then head, then batch
        assert order == [(x, h, b) for b in range(nb) for h in range(ny) for x in range(nx)]
    else:
        body = (ny * nb - k) * nx
        # the body is head-major, the tail block-major over the last k (batch, head) pairs, block 0 (the heaviest) first
        assert order[:body] == [(x, bh % ny, bh // ny) for bh in range(ny * nb - k) for x in range(nx)]
        assert order[body:] == [(x, bh % ny, bh // ny) for x in range(nx) for bh in range(ny * nb - k, ny * nb)]


@pytest.mark.parametrize("nx,ny,nz", [(32, 32, 4), (16, 32, 4), (1, 1, 1), (8, 3, 5), (64, 32, 2), (5, 7, 1), (16, 16, 1)])
@pytest.mark.parametrize("tail_heads", [0, 1, 2, 3, 12, 24, 32, 1000])
def test_grid_kernel_order_is_a_permutation_of_the_grid(nx, ny, nz, tail_heads):
    """The forward's form (blockIdx -> item, shifts and masks only): the last 2^k heads of the LAST batch element are
    dealt block-major, everything before them is untouched."""
    lib = probe_lib()
    l = lib.fa2_probe_tail_log2_for(tail_heads, ny)
    k = 1 << l
    assert (l == 0 and min(tail_heads, ny) < 2) or (k <= min(tail_heads, ny) < 2 * k)
    out = (ctypes.c_int * 3)()
    order = []
    for z in range(nz):          # launch order: x fastest, then y, then z
        for y in range(ny):
            for x in range(nx):
                lib.fa2_probe_tail_block(x, y, z, nx, ny, nz, l, out)
                order.append((out[0], out[1], out[2]))
    assert len(set(order)) == nx * ny * nz and all(0 <= a < nx and 0 <= b < ny and 0 <= c < nz for a, b, c in order)
    plain = [(x, y, z) for z in range(nz) for y in range(ny) for x in range(nx)]
    if l == 0:
        assert order == plain
    else:
        body = (ny * nz - k) * nx
        assert order[:body] == plain[:body]
        assert order[body:] == [(x, ny - k + j, nz - 1) for x in range(nx) for j in range(k)]


def test_tail_length_keeps_the_streamed_operands_in_l2():
    lib = probe_lib()
    assert lib.fa2_probe_tail_heads_for(2 << 20, 0) == 0          # no causal mask: all blocks of a head are equally long
    assert lib.fa2_probe_tail_heads_for(2 << 20, 1) == 24         # config 3: K + V of a head = 2 MB -> 24 heads in 48 MB
    assert lib.fa2_probe_tail_heads_for(64 << 10, 1) == 32        # capped
    assert lib.fa2_probe_tail_heads_for(40 << 20, 1) == 0         # a single head fills the budget: plain order
