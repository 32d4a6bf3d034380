"""Host <-> device copy bandwidth with EVERY GPU of the job copying at the same time (torchrun, one rank per GPU):
512 MiB pinned -> device and device -> pinned per rank, both directions at once - the traffic one `e2e` step of bench.py
moves at config 3. Rank 0 prints the per-rank time and the aggregate, next to the same measurement with one rank alone.
Shows whether the N GPUs of a box share the host's PCIe / memory bandwidth (why `e2e` does not scale with N).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_bw_multi.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def timed(active, reps=5):
    if active:
        both()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            both()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


t_all = timed(True)
t_one = timed(rank == 0)
if rank == 0:
    print(f"{world} GPU(s) copying together: {t_all * 1e3:7.2f} ms per 512 MiB each way and GPU = "
          f"{world * n / t_all / 1e9:6.1f} GB/s aggregate per direction ({n / t_all / 1e9:5.1f} per GPU)")
    print(f"rank 0 alone                : {t_one * 1e3:7.2f} ms = {n / t_one / 1e9:6.1f} GB/s per direction")
if world > 1:
    dist.destroy_process_group()
