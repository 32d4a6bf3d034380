#!/usr/bin/env python
"""Benchmark of the FlashAttention-2 hot path on B200 (contract: see README / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4|cfg5]

One "step" = one pass of the hot path over one batch of synthetic tensors (seed 0, Q/K/V ~ N(0, 0.5^2),
dO ~ N(0, 1): the recipe of the reference's tests/utils.py:21-25). Default workload = BASELINE.json configs[2]
(the configuration its metric is quoted on): bf16 fwd+bwd, B=4 H=32 S=4096 D=128, causal, deterministic dQ, PER GPU
(weak scaling: the global batch is 4*N, sharded by batch with no collective on the data path).

Rank 0 prints ONE JSON line: metric/value (device-timed, inputs resident in HBM), e2e (host buffers in, host
buffers out, through fa2_triton_b200.host_api), roofline of the dominant kernel (per-kernel times taken interleaved in step
order), a `sustained` block (the same step loop for >= 2 s, own clock samples, against the sustained cuBLAS peak),
cpu_baseline, clocks, `triton_baseline` (the reference's own Triton kernels on this GPU when baseline/_ref holds the
reference tree) and, for N > 1, `sharded_matches_single` and a `strong_cfg5` record (BASELINE config 5 partitioned over
the N GPUs).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (B per GPU or global (see "shard"), H, S, D, causal, backward, shard mode)
    "cfg2": dict(B=4, H=32, S=4096, D=128, causal=True, bwd=False, scaling="weak"),
    "cfg3": dict(B=4, H=32, S=4096, D=128, causal=True, bwd=True, scaling="weak"),
    "cfg4": dict(B=2, H=16, S=16384, D=64, causal=False, bwd=True, scaling="weak"),
    "cfg5": dict(B=32, H=32, S=8192, D=128, causal=True, bwd=True, scaling="strong"),
}
METRIC = "fwd+bwd TFLOP/s (bf16, causal, S=4096, D=128)"  # BASELINE.json's metric: the default workload (cfg3)


def metric_name(w):
    """The metric label of a workload: BASELINE.json's for cfg3, the same wording with the workload's own shape otherwise."""
    return (f"{'fwd+bwd' if w['bwd'] else 'fwd'} TFLOP/s (bf16, {'causal' if w['causal'] else 'non-causal'}, "
            f"S={w['S']}, D={w['D']})")


def fa_flops(B, H, Sq, Sk, D, causal, bwd):
    """Algorithmic FLOPs: fwd = 4 B H Sq Sk D (x 1/2 causal), bwd = 2.5 x fwd."""
    f = 4.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
    return f * (3.5 if bwd else 1.0)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.power_w, self.reasons, self.max_mhz = [], [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # NVML missing: clocks are reported as unknown, never invented
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    self.power_w.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                except Exception:
                    pass
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2], "sm_min_mhz": s[0], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(s)}
        if self.power_w:
            pw = sorted(self.power_w)
            out["power_w_median"], out["power_w_max"] = pw[len(pw) // 2], pw[-1]
        return out


def make_inputs(B, H, S, D, device, seed, torch):
    """Q, K, V ~ N(0, 0.5^2), dO ~ N(0, 1) in bf16, generated on the device from a seeded generator."""
    g = torch.Generator(device=device).manual_seed(seed)
    shape = (B, S, H, D)
    out = []
    for std in (0.5, 0.5, 0.5, 1.0):
        t = torch.empty(shape, dtype=torch.bfloat16, device=device)
        for b in range(B):  # per batch element: keeps the fp32 staging small
            t[b] = torch.empty(shape[1:], dtype=torch.float32, device=device).normal_(0, std, generator=g).to(torch.bfloat16)
        out.append(t)
    return out


def cpu_baseline_run(torch, w, steps, warmup, budget_s=20.0):
    """Times the eager-PyTorch port of the reference's CPU path (oracle/eager_reference.py) on a bounded (b, h)
    sample of the workload; the oracle's cost is exactly linear in B*H so the rate carries over."""
    from oracle import eager_reference as er

    # torch.distributed.run exports OMP_NUM_THREADS=1: take every host core explicitly, so that the CPU arm is the same
    # at every --gpus N (only rank 0 runs it; the other ranks exit or idle)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    threads = torch.get_num_threads()
    S, D, causal, bwd = w["S"], w["D"], w["causal"], w["bwd"]
    S_s = min(S, 4096)  # keep the S x S fp32 score tensors within a few hundred MB

    def one(hs):
        g = torch.Generator().manual_seed(0)
        q, k, v = (torch.empty(1, S_s, hs, D).normal_(0, 0.5, generator=g).to(torch.bfloat16) for _ in range(3))
        do = torch.empty(1, S_s, hs, D).normal_(0, 1.0, generator=g).to(torch.bfloat16)
        t0 = time.perf_counter()
        if bwd:
            er.fwd_bwd(q, k, v, do, causal=causal)
        else:
            er.flash_attn_reference(q, k, v, causal=causal)
        return time.perf_counter() - t0

    one(1)
    t1 = one(1)
    total = max(1, steps + warmup)
    hs = int(max(1, min(8, budget_s / total / max(t1, 1e-3))))
    for _ in range(warmup):
        one(hs)
    t0 = time.perf_counter()
    for _ in range(steps):
        one(hs)
    dt = (time.perf_counter() - t0) / max(1, steps)
    flops = fa_flops(1, hs, S_s, S_s, D, causal, bwd)
    return dict(value=flops / dt / 1e12, unit="TFLOP/s", cores=threads, kind="port",
                sample=f"B=1,H={hs},S={S_s},D={D},causal={causal},{'fwd+bwd' if bwd else 'fwd'}; "
                       f"{steps} steps of {dt * 1e3:.0f} ms; eager PyTorch fp32 math on bf16 inputs; "
                       f"torch.set_num_threads(os.cpu_count()={os.cpu_count()})"), dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    w = WORKLOADS[args.workload]
    cb, dt = cpu_baseline_run(torch, w, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": metric_name(w), "value": cb["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, w), "timed_sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(key, w):
    return (f"{key}: bf16 {'fwd+bwd' if w['bwd'] else 'fwd'} B={w['B']},H={w['H']},S={w['S']},D={w['D']} "
            f"{'causal' if w['causal'] else 'non-causal'}" + (", deterministic dQ" if w["bwd"] else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--atomic", action="store_true", help="dQ in arrival order (fast mode) instead of deterministic")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-triton", action="store_true", help="skip the reference's Triton kernels (baseline/_ref)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the config-5 strong-scaling record")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of the `sustained` block (0 = off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import fa2_triton_b200 as fa2
    from fa2_triton_b200 import _lib, shard
    from fa2_triton_b200.host_api import HostAttention

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device - this path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    fa2.set_deterministic(not args.atomic)

    w = dict(WORKLOADS[args.workload])
    H, S, D, causal, bwd = w["H"], w["S"], w["D"], w["causal"], w["bwd"]
    if w["scaling"] == "weak":
        B_local, B_global = w["B"], w["B"] * world
        seed = 0 + rank  # every rank owns different batch elements of the global batch
    else:  # strong: the global (b, h) units are partitioned; this rank materialises only its own slices
        B_global = w["B"]
        blocks = shard.merge_spans(shard.partition(B_global, H, world, rank), H)
        assert len(blocks) == 1 and blocks[0][2] == 0 and blocks[0][3] == H, "cfg5 shards by whole batch elements"
        B_local = blocks[0][1] - blocks[0][0]
        seed = 0 + rank

    q, k, v, do = make_inputs(B_local, H, S, D, dev, seed, torch)
    flops_local = fa_flops(B_local, H, S, S, D, causal, bwd)
    flops_global = fa_flops(B_global, H, S, S, D, causal, bwd)

    def step():
        if bwd:
            qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
            out = fa2.flash_attn_func(qq, kk, vv, causal=causal)
            return torch.autograd.grad(out, (qq, kk, vv), do)
        return fa2.flash_attn_func(q, k, v, causal=causal)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()

    # ---------------------------------------------- timed region ------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = flops_global / (ms_step * 1e-3) / 1e12
    pk = peaks()

    # ------------------------------- sustained: the same step loop for >= 2 s, own clock samples ------------------
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / ms_step)))
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        torch.cuda.synchronize()
        barrier()
        sus_clocks = sampler.stop()
        ts = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sus_ms = ts.item() / n_sus
        sus_val = flops_global / (sus_ms * 1e-3) / 1e12
        sustained = {"value": sus_val, "unit": "TFLOP/s", "steps": n_sus, "ms_per_step": sus_ms, "seconds": ts.item() * 1e-3,
                     "frac_of_measured_sustained_peak": sus_val / world / pk["sustained"],
                     "frac_of_measured_burst_peak": sus_val / world / pk["burst"], "clocks": sus_clocks}

    # -------------------------------- per-kernel timing (roofline of the dominant kernel) ------------------------
    # The step's kernels are launched through the C ABI in STEP ORDER (fwd, delta, dK/dV, dQ, fwd, ...), back to back, one
    # CUDA-event pair per launch on the launching stream and a single synchronize at the end: every kernel is timed in the
    # power / cache state it has inside a step, and the per-kernel times add up to the step (minus launch gaps).
    n_rep = min(args.steps, 20)
    o, lse, scale, _ = fa2._flash_attn_forward(q, k, v, None, None, 0.0, causal, None, None)
    stream = torch.cuda.current_stream().cuda_stream
    S_pad = math.ceil(S / 128) * 128
    s3 = _lib.strides3
    o2, lse2 = torch.empty_like(o), torch.empty_like(lse)
    delta = torch.empty_like(lse)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws_bytes = lib.fa2_sm100_bwd_workspace_bytes(B_local, H, S, S, D, int(causal),
                                                 _lib.DQ_ATOMIC if args.atomic else _lib.DQ_DET_SPILL)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)

    def call_fwd():
        _lib.check(lib.fa2_sm100_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o2.data_ptr(), lse2.data_ptr(),
                                     B_local, H, H, S, S, D, s3(q), s3(k), s3(v), s3(o2), S_pad, scale, int(causal), 17,
                                     stream), "fwd")

    def call_delta():
        _lib.check(lib.fa2_sm100_bwd_delta(o.data_ptr(), do.data_ptr(), delta.data_ptr(), B_local, H, S, D, s3(o), s3(do),
                                           S_pad, 17, ws.data_ptr() if args.atomic else None,
                                           ws_bytes if args.atomic else 0, stream), "delta")

    def call_bwd(mode):
        _lib.check(lib.fa2_sm100_bwd(do.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr(),
                                     delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                     ws.data_ptr(), ws.numel(),
                                     B_local, H, H, S, S, D, s3(do), s3(q), s3(k), s3(v), s3(dq), s3(dk), s3(dv),
                                     S_pad, scale, int(causal), mode, 17, 17, stream), "bwd")

    f_fwd = fa_flops(B_local, H, S, S, D, causal, False)
    seq = [("fa2_fwd2_kernel", call_fwd, f_fwd, None)]
    if bwd:
        seq.append(("fa2_delta_kernel", call_delta, 0.0, None))
        if args.atomic:
            dom = "fa2_bwd_kernel<fused dQ> + fa2_dq_convert_kernel"
            seq.append((dom, lambda: call_bwd(_lib.DQ_ATOMIC), 2.5 * f_fwd, None))
        elif D <= 64:
            # head dim 64: the deterministic backward recomputes S and dP in its dQ kernel instead of spilling dS
            # (fa2_triton_b200/backward.py); algorithmic FLOPs stay 4 + 1 products, the dQ kernel executes 3
            dom = "fa2_bwd_dkdv_kernel"
            seq.append((dom, lambda: call_bwd(_lib.DQ_NONE), 2.0 * f_fwd, None))
            seq.append(("fa2_bwd_dq_kernel (recomputing)", lambda: call_bwd(_lib.DQ_ONLY), 0.5 * f_fwd, None))
        else:
            # S, dP, dV, dK are the dK/dV kernel's products (it also spills dS); the dQ kernel's one product is dS K
            dom = "fa2_bwd_dkdv_kernel<spill>"
            seq.append((dom, lambda: call_bwd(_lib.DQ_SPILL_DKDV), 2.0 * f_fwd, None))
            seq.append(("fa2_bwd_dq_spill_kernel", lambda: call_bwd(_lib.DQ_SPILL_DQ), 0.5 * f_fwd,
                        float(ws_bytes - 1024 + 2 * q.numel() * 2)))
    else:
        dom = "fa2_fwd2_kernel"
    events = {name: [] for name, _, _, _ in seq}
    for i in range(n_rep + 2):
        for name, fn, _, _ in seq:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            events[name].append((a, b))
    torch.cuda.synchronize()
    kern = {}
    for name, _, flop, nbytes in seq:
        ms = sum(a.elapsed_time(b) for a, b in events[name][2:]) / n_rep
        kern[name] = {"ms": ms, "algorithmic_flop": flop, "tflops": flop / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}
        if nbytes is not None:  # HBM-bound kernels: bytes each launch has to move / its duration
            kern[name].update(bound="hbm", algorithmic_bytes=nbytes, gbs=nbytes / (ms * 1e-3) / 1e9,
                              frac_of_hbm_peak=nbytes / (ms * 1e-3) / 1e9 / pk["hbm"])
    achieved = kern[dom]["tflops"]
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["sustained"], "traffic": traffic,
                "peak_source": pk["source"] + " cuBLAS bf16 SUSTAINED (the kernel is timed inside back-to-back steps)",
                "frac_of_burst": achieved / pk["burst"], "frac_of_nominal_2250": achieved / 2250.0,
                "algorithmic_flop_per_launch": kern[dom]["algorithmic_flop"], "kernel_ms": kern[dom]["ms"],
                "sum_of_kernels_ms": sum(v_["ms"] for v_ in kern.values()),
                "timing": "interleaved in step order, one CUDA-event pair per launch", "kernels": kern}

    # ------------------------------------------------- e2e --------------------------------------------------------
    e2e = None
    if bwd and not args.no_e2e:
        plan = HostAttention(q.shape, k.shape, torch.bfloat16, dev)
        hp = [t.cpu().pin_memory() for t in (q, k, v, do)]
        for _ in range(2):
            plan(*hp, causal=causal)
        n_e2e = max(3, min(args.steps, 10))
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # two steps in flight (HostAttention.submit / HostTicket.result): every step still uploads its own inputs and
        # downloads its own results inside the timed region; the next step's upload overlaps this step's download
        prev = None
        for _ in range(n_e2e):
            ticket = plan.submit(*hp, causal=causal)
            if prev is not None:
                _ = float(prev.result()[0][0, 0, 0, 0])  # the host really holds the result
            prev = ticket
        _ = float(prev.result()[0][0, 0, 0, 0])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": flops_global / tt.item() / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": plan.h2d_bytes,
               "d2h_bytes_per_step": plan.d2h_bytes, "ms_per_step": tt.item() * 1e3, "steps": n_e2e,
               "api": "fa2_triton_b200.host_api.HostAttention (pinned host q,k,v,dO in; o,dq,dk,dv out; "
                      f"{plan.chunks} (batch, head range) chunks over 3 streams, pitched copies, {plan.depth} steps in flight)",
               "bound": "host<->device link: 512 MiB each way per step and GPU at config 3; the N GPUs of a box share the "
                        "host's PCIe / memory bandwidth, so e2e does not scale with N (scripts/pcie_bw.py)"}
        del plan, hp
        torch.cuda.empty_cache()

    # ---------------------- sharded == single: a foreign shard recomputed on rank 0 (verification only) -----------
    sharded_ok = None
    if world > 1 and bwd:
        def checksums(ts):
            # exact and order-independent: the 16-bit patterns summed as integers
            return [int(t_.contiguous().view(torch.int16).to(torch.int64).sum().item()) for t_ in ts]

        mine = checksums((step_out := run_step_outputs(fa2, torch, q, k, v, do, causal)))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            other = 1  # rank 1's shard: weak scaling = its own batch elements (seed 1), strong = the next batch range
            if w["scaling"] == "weak":
                fq, fk, fv, fdo = make_inputs(B_local, H, S, D, dev, 0 + other, torch)
            else:
                ob = shard.merge_spans(shard.partition(B_global, H, world, other), H)[0]
                fq, fk, fv, fdo = make_inputs(ob[1] - ob[0], H, S, D, dev, 0 + other, torch)
            sharded_ok = checksums(run_step_outputs(fa2, torch, fq, fk, fv, fdo, causal)) == gathered[other]
            del fq, fk, fv, fdo
        del step_out

    # ------------------------------- strong scaling of BASELINE config 5 over the N GPUs --------------------------
    strong = None
    if world > 1 and args.workload == "cfg3" and not args.no_strong:
        del q, k, v, do, o, o2, dq, dk, dv, ws
        torch.cuda.empty_cache()
        strong = strong_cfg5(fa2, shard, torch, dist, dev, rank, world, barrier)

    # -------------------------------------------- CPU baseline (rank 0, N=1) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_run(torch, w, steps=3, warmup=1, budget_s=20.0)

    # ------------------------ the reference's own Triton kernels on this GPU (rank 0, needs baseline/_ref) ---------
    triton_base = None
    if rank == 0 and not args.no_triton and args.workload in ("cfg2", "cfg3"):
        triton_base = triton_baseline_run(torch, w, dev, ms_step if world == 1 else None)
    barrier()

    if rank == 0:
        line = {
            "metric": metric_name(w), "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, w), "per_gpu_batch": B_local, "global_batch": B_global,
                       "sharding": f"batch x head units over {world} GPU(s), no data-path collective",
                       "dq_mode": "atomic" if args.atomic else "deterministic (dS tiles spilled by the dK/dV kernel, dQ = sum dS K in fixed order)",
                       "l2": "inputs (q,k,v,o,dO = 5 x 128 MiB at cfg3) exceed the 126 MB L2; no explicit flush",
                       "api": "fa2_triton_b200.flash_attn_func + torch.autograd.grad"},
            "frac_of_measured_burst_peak": value / world / pk["burst"],
            "frac_of_measured_sustained_peak": value / world / pk["sustained"],
            "frac_of_nominal_peak_2250": value / world / 2250.0,
            "sustained": sustained, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "triton_baseline": triton_base,
            "gpu_launches": args.steps * (4 if bwd else 1),  # fwd, delta, dK/dV (+dS tiles), dQ kernel or dQ convert
        }
        if world > 1:
            line["sharded_matches_single"] = sharded_ok
            line["strong_cfg5"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_step_outputs(fa2, torch, q, k, v, do, causal):
    qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=causal)
    grads = torch.autograd.grad(out, (qq, kk, vv), do)
    torch.cuda.synchronize()
    return (out.detach(),) + tuple(grads)


def strong_cfg5(fa2, shard, torch, dist, dev, rank, world, barrier, steps=5):
    """BASELINE config 5 (B=32, H=32, S=8192, D=128 causal fwd+bwd) partitioned over the `world` GPUs by
    shard.partition (whole batch elements per rank): total work fixed, so value(N) / value(1) is the strong-scaling
    speed-up. Timed like the main region: CUDA events, max over ranks."""
    w = WORKLOADS["cfg5"]
    B, H, S, D = w["B"], w["H"], w["S"], w["D"]
    blocks = shard.merge_spans(shard.partition(B, H, world, rank), H)
    if len(blocks) != 1 or blocks[0][2] != 0 or blocks[0][3] != H:
        return {"unavailable": f"{world} ranks do not cut B={B} into whole batch elements"}
    B_local = blocks[0][1] - blocks[0][0]
    q, k, v, do = make_inputs(B_local, H, S, D, dev, 100 + rank, torch)

    def step():
        qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, causal=True)
        return torch.autograd.grad(out, (qq, kk, vv), do)

    for _ in range(3):
        step()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    flops = fa_flops(B, H, S, S, D, True, True)
    return {"workload": workload_name("cfg5", w), "scaling": "strong", "n_gpus": world, "per_gpu_batch": B_local,
            "steps": steps, "warmup": 3, "ms_per_step": ms, "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
            "note": "value at N=1 (whole problem on one GPU): python bench.py --workload cfg5"}


def triton_baseline_run(torch, w, dev, ours_ms, steps=10):
    """The reference's own Triton kernels (src/wrapper.py:89-100, UNMODIFIED, imported from the git-ignored copy of the
    reference tree in baseline/_ref) on this GPU, same workload, same inputs recipe, CUDA events around back-to-back
    steps after an autotuning first call + 3 warm-up steps. Returns None when baseline/_ref is absent."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "src" / "wrapper.py").exists():
        return None
    try:
        import types

        import triton.language as tl
        if not hasattr(tl, "const_pointer_type"):  # the reference predates Triton 3.6
            tl.const_pointer_type = tl.pointer_type
        tmod = types.ModuleType("tests")          # another installed package is called `tests`
        tmod.__path__ = [str(ref / "tests")]
        sys.modules.setdefault("tests", tmod)
        if str(ref) not in sys.path:
            sys.path.insert(0, str(ref))
        from src.wrapper import flash_attn_func as ref_func

        B, H, S, D, causal, bwd = w["B"], w["H"], w["S"], w["D"], w["causal"], w["bwd"]
        q, k, v, do = make_inputs(B, H, S, D, dev, 0, torch)

        def step():
            if bwd:
                qq, kk, vv = (t.detach().requires_grad_(True) for t in (q, k, v))
                out = ref_func(qq, kk, vv, causal=causal)
                return torch.autograd.grad(out, (qq, kk, vv), do)
            return ref_func(q, k, v, causal=causal)

        for _ in range(4):  # the first call autotunes
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"value": fa_flops(B, H, S, S, D, causal, bwd) / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms,
               "steps": steps, "what": "remi-or/fa2_triton src.wrapper.flash_attn_func (Triton 3.6, autotuned), unmodified, "
                                       "one GPU, per-GPU workload"}
        if ours_ms:
            out["speedup_ours_over_triton"] = ms / ours_ms
        return out
    except Exception as e:  # a Triton that cannot compile / launch the reference is also a result
        return {"unavailable": repr(e)[:500]}


if __name__ == "__main__":
    main()
