"""Which reading of LBO / SBO does the no-swizzle MN-major A descriptor want? (a_mode 3 vs 4 of the UMMA probe)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fa2_triton_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
for a_mode in (3, 4, 1):
    for n, k in ((128, 128), (64, 64)):
        g = torch.Generator(device="cpu").manual_seed(1)
        a = torch.randn(128, k, generator=g).bfloat16().to(dev)
        b = torch.randn(k, n, generator=g).bfloat16().to(dev)
        ref = a.float() @ b.float()
        d = torch.full((128, n), float("nan"), device=dev)
        code = lib.fa2_sm100_debug_umma_probe(a.t().contiguous().data_ptr(), b.contiguous().data_ptr(), d.data_ptr(), n, k, a_mode, 1,
                                              _lib.DTYPE_BF16, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        print(f"a_mode={a_mode} N={n} K={k} rc={code} max|err|={(d - ref).abs().max().item():.4g}")
