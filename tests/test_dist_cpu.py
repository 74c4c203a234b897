"""CPU, world size 2 over gloo: the host-side multi-GPU logic (gradient bucket all-reduce driven by the
network's grad_ready_hook protocol, volume sharding + final gather/sort of the FPL uncertainties)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fplplus_b200 import agent as A, fpl
        # ---- gradient reducer: the hook protocol of UNet2D5_dsbn._run_backward (flat, start, end, last) ----
        g = torch.Generator().manual_seed(100 + rank)
        n = 3_000_000                                        # 12 MB of fp32 "gradients"
        flat = torch.randn(n, generator=g)
        mine = flat.clone()
        red = A.GradAllReducer(bucket_bytes=4 << 20)
        cuts = [0, 10, 700_000, 700_100, 1_900_000, 2_999_990, n]       # irregular completion order
        for i in range(len(cuts) - 1):
            red.hook(flat, cuts[i], cuts[i + 1], i == len(cuts) - 2)
        n_buckets = len(red._pending)
        red.finish()
        others = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(others, mine)
        ref = torch.stack(others).mean(0)
        ok_avg = bool(torch.allclose(flat, ref, rtol=1e-6, atol=1e-7))
        # ---- the second domain pass of the same step reuses the reducer after finish() (net.grad_wait_hook): the
        #      network hands over >= grad_bucket_bytes at a time, the last range always ----
        flat2 = torch.randn(n, generator=g)
        mine2 = flat2.clone()
        for a, b, last in ((0, 1_100_000, False), (1_100_000, 2_400_000, False), (2_400_000, n, True)):
            red.hook(flat2, a, b, last)
        red.finish()
        assert not red._pending and red._start is None
        others2 = [torch.empty_like(mine2) for _ in range(world)]
        dist.all_gather(others2, mine2)
        ok_avg = ok_avg and bool(torch.allclose(flat2, torch.stack(others2).mean(0), rtol=1e-6, atol=1e-7))
        # ---- deferred mode: ONE all-reduce of the master gradient buffer behind every p.grad (both domain passes already
        #      summed into it), and the per-parameter fallback when the gradients came through autograd ----
        class _Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a = torch.nn.Parameter(torch.zeros(5, 3))
                self.b = torch.nn.Parameter(torch.zeros(7))
                self.frozen = torch.nn.Parameter(torch.zeros(2))            # never receives a gradient
        net = _Net()
        buf = torch.randn(24, generator=g)
        mine3 = buf.clone()
        net._master = {"buf": buf}
        net.a.grad, net.b.grad = buf[0:15].view(5, 3), buf[16:23]
        A.DeferredGradAllReducer(net).finish()
        others3 = [torch.empty_like(mine3) for _ in range(world)]
        dist.all_gather(others3, mine3)
        ok_avg = ok_avg and bool(torch.allclose(buf, torch.stack(others3).mean(0), rtol=1e-6, atol=1e-7))
        ok_avg = ok_avg and net.a.grad.data_ptr() == buf.data_ptr() and net.frozen.grad is None
        net._master = None
        net.a.grad, net.b.grad = mine3[0:15].clone().view(5, 3), mine3[16:23].clone()
        A.DeferredGradAllReducer(net).finish()
        ok_avg = ok_avg and bool(torch.allclose(net.b.grad, buf[16:23], rtol=1e-6, atol=1e-7))
        # ---- inference: volumes round-robin, gather of (value, name) pairs, host sort ----
        cfg = {"dataset": {"tensor_type": "float"}, "network": {}, "training": {}, "testing": {}}
        ag = A.SegmentationAgent(cfg, "test")
        names = ["vol_%02d.nii.gz" % i for i in range(7)]
        table = {nm: [1 if i == 3 else 0.01 * (7 - i)] for i, nm in enumerate(names)}
        local = {nm: table[nm] for nm in A.shard_round_robin(names, ag.rank, ag.world)}
        merged = ag._gather_dict(local)
        srt = fpl.sort_uncertainty(merged)
        # ---- rank 0's parameters / BatchNorm buffers and validation scalars on every rank ----
        torch.manual_seed(7 + rank)                          # ranks start from DIFFERENT weights and running statistics
        ag.device = torch.device("cpu")
        ag.net = torch.nn.Sequential(torch.nn.Conv3d(1, 4, 3), torch.nn.BatchNorm3d(4))
        ag.net[1].running_mean.normal_()
        ag.net[1].num_batches_tracked.fill_(5 + rank)
        ag._broadcast_state()
        state = torch.cat([t.detach().double().reshape(-1) for t in list(ag.net.parameters()) + list(ag.net.buffers())])
        all_states = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(all_states, state)
        ok_bcast = all(torch.equal(s_, all_states[0]) for s_ in all_states) and int(ag.net[1].num_batches_tracked) == 5
        ag.net[1].running_var.fill_(1.0 + rank)              # rank-local BN statistics diverge again during training
        ag._broadcast_state(buffers_only=True)
        ok_bcast = ok_bcast and float(ag.net[1].running_var[0]) == 1.0
        scal = ag._agree_scalars({"loss": 0.5 + rank, "avg_dice": 0.25 * (rank + 1), "class_dice": np.array([0.1, 0.4]) * (rank + 1)})
        ok_bcast = ok_bcast and scal["loss"] == 0.5 and scal["avg_dice"] == 0.25 and np.allclose(scal["class_dice"], [0.1, 0.4])
        q.put((rank, ok_avg and ok_bcast, n_buckets, len(local), [nm for _v, nm in srt], ag.world))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_and_sharded_inference():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    names = ["vol_%02d.nii.gz" % i for i in range(7)]
    expected = [names[i] for i in (6, 5, 4, 2, 1, 0, 3)]        # ascending uncertainty, the sentinel (1) last
    for rank, ok_avg, n_buckets, n_local, order, w in res:
        assert ok_avg, "rank %d: all-reduced buckets != mean over ranks" % rank
        assert n_buckets == 3                                    # 12 MB in >= 4 MB buckets + the tail flush
        assert w == 2 and n_local == (4 if rank == 0 else 3)
        assert order == expected
