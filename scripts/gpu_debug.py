"""Bring-up driver for a GPU box: runs each stage in its own process (so a trap in one does not hide the others)
and prints compact diagnostics. Usage: python scripts/gpu_debug.py [probe] [fwd] [bwd]"""
import subprocess
import sys

STAGES = {
    "probe": r"""
import torch
from fa2_triton_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
for (am, bm) in [(0,0),(0,1),(2,1),(2,0),(1,1),(1,0)]:
    for (n, k) in [(128,128),(64,128),(128,64),(64,64)]:
        g = torch.Generator().manual_seed(1)
        a = torch.randn(128, k, generator=g).bfloat16().to(dev)
        b = torch.randn(k, n, generator=g).bfloat16().to(dev)
        ref = a.float() @ b.float()
        a_arg = a.t().contiguous() if am == 1 else a.contiguous()
        b_arg = b.t().contiguous() if bm == 0 else b.contiguous()
        d = torch.full((128, n), float("nan"), device=dev)
        rc = lib.fa2_sm100_debug_umma_probe(a_arg.data_ptr(), b_arg.data_ptr(), d.data_ptr(), n, k, am, bm, 17, None)
        torch.cuda.synchronize()
        err = (d - ref).abs().max().item()
        print(f"probe a_mode={am} b_mode={bm} N={n} K={k}: rc={rc} maxerr={err:.4g} {'OK' if err < 1e-2 else 'FAIL'}", flush=True)
""",
    "fwd": r"""
import torch, math
import fa2_triton_b200 as fa2
from oracle import eager_reference as er
dev = "cuda:0"
for (B,Sq,Sk,H,D,causal) in [(1,128,128,1,128,False),(1,256,256,1,128,False),(1,256,256,2,128,True),(2,512,512,3,128,True),
                             (1,128,128,1,64,False),(2,512,512,3,64,True),(2,1000,1000,2,128,True),(1,113,203,2,64,False),(1,4096,4096,2,128,True)]:
    torch.manual_seed(0)
    q = torch.empty(B,Sq,H,D).normal_(0,0.5).bfloat16().to(dev)
    k = torch.empty(B,Sk,H,D).normal_(0,0.5).bfloat16().to(dev)
    v = torch.empty(B,Sk,H,D).normal_(0,0.5).bfloat16().to(dev)
    o, lse, _, _ = fa2._flash_attn_forward(q,k,v,None,None,0.0,causal,None,None)
    torch.cuda.synchronize()
    ref = er.flash_attn_reference(q,k,v,causal=causal)
    lse_ref = er.lse_log2(q,k,causal=causal)
    eo = (o.float()-ref.float()).abs().max().item()
    el = (lse[:,:,:Sq]-lse_ref).abs().max().item()
    print(f"fwd B={B} Sq={Sq} Sk={Sk} H={H} D={D} causal={causal}: O maxerr={eo:.4g} LSE maxerr={el:.4g} {'OK' if eo < 2e-2 and el < 1e-3 else 'FAIL'}", flush=True)
""",
    "bwd": r"""
import torch, math
import fa2_triton_b200 as fa2
from oracle import eager_reference as er
dev = "cuda:0"
for det in (True, False):
  for (B,Sq,Sk,H,D,causal) in [(1,128,128,1,128,False),(1,256,256,1,128,False),(1,256,256,2,128,True),(2,512,512,3,128,True),
                             (1,128,128,1,64,False),(2,512,512,3,64,True),(2,1000,1000,2,128,True),(1,113,203,2,64,False),(1,2048,2048,2,128,True)]:
    torch.manual_seed(0)
    q = torch.empty(B,Sq,H,D).normal_(0,0.5).bfloat16().to(dev)
    k = torch.empty(B,Sk,H,D).normal_(0,0.5).bfloat16().to(dev)
    v = torch.empty(B,Sk,H,D).normal_(0,0.5).bfloat16().to(dev)
    do = torch.randn(B,Sq,H,D).bfloat16().to(dev)
    fa2.set_deterministic(det)
    qq,kk,vv = (t.clone().requires_grad_(True) for t in (q,k,v))
    out = fa2.flash_attn_func(qq,kk,vv,causal=causal)
    dq,dk,dv = torch.autograd.grad(out,(qq,kk,vv),do)
    torch.cuda.synchronize()
    r = er.fwd_bwd(q,k,v,do,causal=causal)
    errs = [(a.float()-b.float()).abs().max().item() for a,b in zip((out,dq,dk,dv), r)]
    mags = [b.float().abs().max().item() for b in r]
    ok = all(e < 2e-2 + 1e-2*m for e,m in zip(errs,mags))
    print(f"bwd det={det} B={B} Sq={Sq} Sk={Sk} H={H} D={D} causal={causal}: O/dQ/dK/dV maxerr=" + "/".join(f"{e:.3g}" for e in errs) + " (max|ref| " + "/".join(f"{m:.3g}" for m in mags) + f") {'OK' if ok else 'FAIL'}", flush=True)
""",
}

if __name__ == "__main__":
    stages = sys.argv[1:] or list(STAGES)
    for s in stages:
        print(f"===== {s} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, "-c", STAGES[s]], timeout=240)
            print(f"===== {s}: exit {r.returncode} =====", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== {s}: TIMEOUT =====", flush=True)
