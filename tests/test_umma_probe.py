"""GPU: pin the UMMA descriptor encodings (ptx.cuh) against a torch matmul, one operand path at a time.
The probe kernel lives in its own test-only library (tests/libfa2_probe.so, built by csrc/Makefile from
csrc/probe_sm100.cu); it is not part of libfa2_sm100.so nor of its header."""
import ctypes
from pathlib import Path

import pytest
import torch

from fa2_triton_b200 import _lib


def probe_lib():
    path = Path(__file__).resolve().parent / "libfa2_probe.so"
    assert path.exists(), f"{path} missing - run `make -C fa2_triton_b200/csrc`"
    lib = ctypes.CDLL(str(path))
    lib.fa2_probe_umma.restype = ctypes.c_int
    lib.fa2_probe_umma.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    return lib

pytestmark = pytest.mark.gpu

MODES = [(0, 0), (0, 1), (2, 1), (2, 0), (1, 1), (1, 0)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,k", [(128, 128), (64, 128), (128, 64), (64, 64)])
@pytest.mark.parametrize("a_mode,b_mode", MODES)
def test_probe_matches_matmul(a_mode, b_mode, n, k, dtype):
    lib = probe_lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(a_mode * 10 + b_mode)
    a = torch.randn(128, k, generator=g).to(dtype).to(dev)   # logical A [M=128, K]
    b = torch.randn(k, n, generator=g).to(dtype).to(dev)     # logical B [K, N]
    ref = a.float() @ b.float()
    a_arg = a.t().contiguous() if a_mode == 1 else a.contiguous()  # mode 1 takes A^T [K, 128]
    b_arg = b.t().contiguous() if b_mode == 0 else b.contiguous()  # mode 0 takes B^T [N, K]
    d = torch.full((128, n), float("nan"), device=dev, dtype=torch.float32)
    code = lib.fa2_probe_umma(a_arg.data_ptr(), b_arg.data_ptr(), d.data_ptr(), n, k, a_mode, b_mode,
                              _lib.DTYPE_BF16 if dtype == torch.bfloat16 else _lib.DTYPE_FP16,
                              torch.cuda.current_stream().cuda_stream)
    assert code == 0, f"fa2_probe_umma returned {code}"
    torch.cuda.synchronize()
    torch.testing.assert_close(d, ref, rtol=1e-4, atol=1e-3)
