"""Multi-GPU check of the batch x head partitioner (run under torchrun, one process per GPU, NCCL):
every rank runs fwd + bwd on its (batch, kv-head) slices only, the outputs are gathered with NCCL for verification,
and rank 0 compares them bit for bit with the same computation done on one GPU. No collective on the data path.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/verify_sharded.py"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
import fa2_triton_b200 as fa2
from fa2_triton_b200 import shard

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S, Hq, Hkv, D = 3, 640, 8, 4, 128
g = torch.Generator(device="cpu").manual_seed(0)  # identical global tensors on every rank; each touches only its slices
q = torch.empty(B, S, Hq, D).normal_(0, 0.5, generator=g).bfloat16().to(dev)
k = torch.empty(B, S, Hkv, D).normal_(0, 0.5, generator=g).bfloat16().to(dev)
v = torch.empty(B, S, Hkv, D).normal_(0, 0.5, generator=g).bfloat16().to(dev)
do = torch.empty(B, S, Hq, D).normal_(0, 1.0, generator=g).bfloat16().to(dev)
outs, dqs, dks, dvs, blocks = [], [], [], [], []
for block, qs, ks, vs, dos in shard.shard_views(q, k, v, world, rank, do):
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (qs, ks, vs))
    o = fa2.flash_attn_func(qq, kk, vv, causal=True)
    dq_, dk_, dv_ = torch.autograd.grad(o, (qq, kk, vv), dos)
    blocks.append(block); outs.append(o.detach()); dqs.append(dq_); dks.append(dk_); dvs.append(dv_)
grp = Hq // Hkv
full = [shard.gather_for_verification(t, shp, blocks, heads_per_unit=hpu, dtype=q.dtype, device=dev)
        for t, shp, hpu in ((outs, q.shape, grp), (dqs, q.shape, grp), (dks, k.shape, 1), (dvs, k.shape, 1))]
if rank == 0:
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    o = fa2.flash_attn_func(qq, kk, vv, causal=True)
    ref = (o.detach(),) + torch.autograd.grad(o, (qq, kk, vv), do)
    ok = all(torch.equal(a, b) for a, b in zip(full, ref))
    print(f"sharded over {world} GPUs == single GPU, bit for bit: {ok}")
    assert ok
dist.barrier(); dist.destroy_process_group()
