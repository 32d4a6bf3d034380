"""GPU: the host-buffer entry point (what bench.py times as e2e) returns exactly what the device API returns."""
import pytest
import torch

import fa2_triton_b200 as fa2
from fa2_test_utils import generate_test_data
from fa2_triton_b200.host_api import HostAttention

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,S,Hq,Hkv,D,chunks", [(2, 640, 8, 8, 128, None), (3, 300, 8, 2, 64, 6), (1, 1000, 4, 4, 128, 3)])
def test_host_api_equals_device_api(B, S, Hq, Hkv, D, chunks):
    q, k, v, do = generate_test_data(B, S, S, Hq, Hkv, D, torch.bfloat16, DEV, seed=7)
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = fa2.flash_attn_func(qq, kk, vv, causal=True)
    dq, dk, dv = torch.autograd.grad(out, (qq, kk, vv), do)
    plan = HostAttention(q.shape, k.shape, torch.bfloat16, DEV, chunks)
    host = [t.cpu().pin_memory() for t in (q, k, v, do)]
    for _ in range(2):  # the plan is reusable
        o_h, dq_h, dk_h, dv_h = plan(*host, causal=True)
        for got, want in ((o_h, out), (dq_h, dq), (dk_h, dk), (dv_h, dv)):
            assert not got.is_cuda and got.is_pinned()
            assert torch.equal(got, want.detach().cpu())


def test_host_api_several_calls_in_flight():
    """submit / result with two calls in flight and slot reuse: different inputs per call, every ticket returns its own
    call's results (compared with the device API), in any order of collection."""
    B, S, H, D = 2, 512, 4, 128
    plan = HostAttention((B, S, H, D), (B, S, H, D), torch.bfloat16, DEV, chunks=4, depth=2)
    want, tickets = [], []
    for step in range(5):
        q, k, v, do = generate_test_data(B, S, S, H, H, D, torch.bfloat16, DEV, seed=40 + step)
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        out = fa2.flash_attn_func(qq, kk, vv, causal=True)
        want.append([t.detach().cpu() for t in (out,) + torch.autograd.grad(out, (qq, kk, vv), do)])
        host = [t.cpu().pin_memory() for t in (q, k, v, do)]
        tickets.append((plan.submit(*host, causal=True), host))
        if step >= 1:  # collect the previous call while this one is in flight
            got = tickets[step - 1][0].result()
            for g, w in zip(got, want[step - 1]):
                assert torch.equal(g, w)
    for g, w in zip(tickets[-1][0].result(), want[-1]):
        assert torch.equal(g, w)
