"""GPU: FusedAdam (csrc/adam.cu, one launch over all tensors) against torch.optim.Adam(weight_decay) -- the optimiser
net_run/get_optimizer.py:16-17 builds -- at rtol 1e-6 over 10 steps, including parameters that miss gradients in some
steps (per-parameter step counters), sizes that are not multiples of 4, a device-side learning rate, and state_dict
round trips in both directions."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(16, 1, 3, 3, 3), (16,), (1,), (32, 16, 3, 3, 3), (7, 5), (256, 128, 3, 3, 3), (3,), (4099,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in SHAPES]


def _grads(step, params, skip):
    g = torch.Generator().manual_seed(1000 + step)
    out = []
    for i, p in enumerate(params):
        gr = (torch.randn(p.shape, generator=g) * (10.0 ** ((i % 3) - 2))).to(DEV)
        out.append(None if i in skip else gr)
    return out


@pytest.mark.parametrize("device_lr", [False, True])
def test_fused_adam_matches_torch_adam(device_lr):
    from fplplus_b200.optim import FusedAdam
    pa, pb = _params(3), _params(3)
    lr = 2e-3
    ours = FusedAdam(pa, torch.tensor(lr, dtype=torch.float32, device=DEV) if device_lr else lr, weight_decay=1e-5)
    ref = torch.optim.Adam(pb, lr, weight_decay=1e-5)
    for step in range(10):
        skip = {2, 6} if step in (1, 2, 5) else set()          # these tensors keep their own step counters
        if step == 6:                                          # MultiStepLR-style change
            lr *= 0.5
            if device_lr:
                ours.param_groups[0]["lr"].fill_(lr)
            else:
                ours.param_groups[0]["lr"] = lr
            ref.param_groups[0]["lr"] = lr
        for p, q, g in zip(pa, pb, _grads(step, pa, skip)):
            p.grad = None if g is None else g.clone()
            q.grad = None if g is None else g.clone()
        ours.step()
        ref.step()
        for i, (p, q) in enumerate(zip(pa, pb)):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-6, atol=1e-7, msg=lambda m: "step %d tensor %d: %s" % (step, i, m))
    for i, (p, q) in enumerate(zip(pa, pb)):
        so, sr = ours.state[p], ref.state[q]
        assert float(so["step"]) == float(sr["step"]) == (7.0 if i in (2, 6) else 10.0)
        torch.testing.assert_close(so["exp_avg"], sr["exp_avg"], rtol=1e-6, atol=1e-9)
        torch.testing.assert_close(so["exp_avg_sq"], sr["exp_avg_sq"], rtol=1e-6, atol=1e-12)


def test_fused_adam_state_dict_round_trips_with_torch_adam():
    from fplplus_b200.optim import FusedAdam
    pa, pb, pc, pd = _params(5), _params(5), _params(5), _params(5)
    ours, ref = FusedAdam(pa, 1e-3, weight_decay=1e-5), torch.optim.Adam(pb, 1e-3, weight_decay=1e-5)
    for step in range(3):
        for p, q, g in zip(pa, pb, _grads(step, pa, set())):
            p.grad, q.grad = g.clone(), g.clone()
        ours.step()
        ref.step()
    sd_ours, sd_ref = copy.deepcopy(ours.state_dict()), copy.deepcopy(ref.state_dict())
    assert set(sd_ours) == set(sd_ref) == {"state", "param_groups"}
    assert set(sd_ours["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    for k in ("lr", "betas", "eps", "weight_decay"):
        assert sd_ours["param_groups"][0][k] == sd_ref["param_groups"][0][k]
    # torch -> ours and ours -> torch, then 3 more steps everywhere: all four trajectories coincide
    with torch.no_grad():
        for src, dst in ((pb, pc), (pa, pd)):
            for s_, d_ in zip(src, dst):
                d_.copy_(s_)
    ours2 = FusedAdam(pc, 1e-3, weight_decay=1e-5)
    ours2.load_state_dict(sd_ref)
    ref2 = torch.optim.Adam(pd, 1e-3, weight_decay=1e-5)
    ref2.load_state_dict(sd_ours)
    for step in range(3, 6):
        gs = _grads(step, pa, set())
        for plist in (pa, pb, pc, pd):
            for p, g in zip(plist, gs):
                p.grad = g.clone()
        for o in (ours, ref, ours2, ref2):
            o.step()
    for a, b, c, d in zip(pa, pb, pc, pd):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(c.detach(), b.detach(), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(d.detach(), b.detach(), rtol=1e-6, atol=1e-7)
    assert float(ours2.state[pc[0]]["step"]) == 6.0


def test_fused_adam_inside_a_cuda_graph_follows_the_device_lr():
    from fplplus_b200.optim import FusedAdam
    pa, pb = _params(9), _params(9)
    lr_t = torch.tensor(1e-3, dtype=torch.float32, device=DEV)
    ours = FusedAdam(pa, lr_t, weight_decay=1e-5)
    ref = torch.optim.Adam(pb, 1e-3, weight_decay=1e-5)
    static_g = [torch.zeros_like(p) for p in pa]
    for p, g in zip(pa, static_g):
        p.grad = g
    ours.step()                                                # warm-up builds the table (allocations outside the capture)
    for q, g in zip(pb, static_g):
        q.grad = g.clone()
    ref.step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    for step in range(4):
        gs = _grads(step, pa, set())
        for sg, q, g in zip(static_g, pb, gs):
            sg.copy_(g)
            q.grad = g.clone()
        if step == 2:
            lr_t.fill_(2.5e-4)
            ref.param_groups[0]["lr"] = 2.5e-4
        graph.replay()
        ref.step()
    for p, q in zip(pa, pb):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-6, atol=1e-7)
    assert float(ours.state[pa[0]]["step"]) == 5.0
