"""torchrun --nproc-per-node 2 tools/ddp_check.py : data-parallel consistency of SegmentationAgent.train_step.
Every rank trains on DIFFERENT synthetic batches (eager warm-up steps, then CUDA-graph replays); after each step the
parameters must be bit-identical on all ranks (the all-reduced gradients, not the local ones, reached the optimiser),
and they must differ from a run without the all-reduce.  Prints one OK line on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench


def checksum(net):
    return torch.stack([p.detach().double().sum() for p in net.parameters()] +
                       [p.detach().double().abs().sum() for p in net.parameters()])


def grads_of(agent, batches):
    """Gradients (flat fp64 copy) of one training_all forward/backward of `batches`, optimiser untouched."""
    agent.optimizer.zero_grad(set_to_none=True)
    total = None
    for d, data in enumerate(batches):
        x, y = agent._to_device(data["image"]), agent._to_device(data["label_prob"])
        out = agent.net(x, domain_label=d * torch.ones(x.shape[0], dtype=torch.long))
        loss_d = agent.get_loss_value(data, out, y, agent.fpl_uda)
        total = loss_d if total is None else total + loss_d
    (total / len(batches)).backward()
    if agent.reducer is not None:
        agent.reducer.finish()
    torch.cuda.synchronize()
    return {n: p.grad.detach().double().clone() for n, p in agent.net.named_parameters() if p.grad is not None}


def grad_mean_check(agent, rank, world):
    """The all-reduced gradient of this rank's step == mean over ranks of the gradients each rank's batch gives in a
    single process (same weights everywhere): what nn.DataParallel's gather-then-backward amounts to per sample
    group (agent_seg.py:695), up to the per-rank Dice normalisation documented in DESIGN.md section 4."""
    small = (16, 64, 64)
    per_rank = [[bench.make_batch(500 + r * 2, 2, small, False, False), bench.make_batch(501 + r * 2, 2, small, True, False)]
                for r in range(world)]
    reducer, hook, wait = agent.reducer, agent.net.grad_ready_hook, agent.net.grad_wait_hook
    drops = [m for m in agent.net.modules() if type(m) == torch.nn.Dropout]
    for m in drops:
        m.eval()                                                     # identical arithmetic in every pass below
    averaged = grads_of(agent, per_rank[rank])                       # with the overlapped NCCL all-reduce
    agent.reducer, agent.net.grad_ready_hook, agent.net.grad_wait_hook = None, None, None
    singles = [grads_of(agent, per_rank[r]) for r in range(world)]   # no communication: every rank computes all of them
    agent.reducer, agent.net.grad_ready_hook, agent.net.grad_wait_hook = reducer, hook, wait
    agent.optimizer.zero_grad(set_to_none=True)
    for m in drops:
        m.train()
    worst = ("", 0.0)
    for k, g in averaged.items():
        ref = sum(s[k] for s in singles) / world
        err = float((g - ref).norm() / (ref.norm() + 1e-30))
        if err > worst[1]:
            worst = (k, err)
    local_vs_avg = float((singles[rank]["out_conv.weight"] - averaged["out_conv.weight"]).norm()
                         / averaged["out_conv.weight"].norm())
    if worst[1] > 1e-3:
        raise SystemExit("rank %d: all-reduced gradient differs from the mean of the per-rank gradients: %s rel %.3e" % (
            (rank,) + worst))
    if local_vs_avg < 1e-3:
        raise SystemExit("rank %d: the averaged gradient equals the local one -- no all-reduce happened" % rank)
    dist.barrier()
    if rank == 0:
        print("GRAD MEAN OK: worst rel_l2 %.2e (%s); local vs averaged %.2e" % (worst[1], worst[0], local_vs_avg), flush=True)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from fplplus_b200.agent import reserve_sms_for_nccl
    reserve_sms_for_nccl()
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    agent = bench.build_agent("train", world)
    if os.environ.get("GRAD_MEAN", "0") != "0":
        grad_mean_check(agent, rank, world)
    exact = os.environ.get("EXACT_DP", "0") != "0"
    if exact:
        agent.config["training"]["exact_dp_dice"] = True
        agent.world = world
        agent._install_exact_dp_dice()
    host = [bench.make_batch(100 + rank * 2, bench.BATCH, bench.PATCH, False, True),
            bench.make_batch(101 + rank * 2, bench.BATCH, bench.PATCH, True, True)]
    steps = int(os.environ.get("STEPS", "7"))
    for it in range(steps):
        loss, _ = agent.train_step(host)
        torch.cuda.synchronize()
        mine = checksum(agent.net)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for r in range(world):
            if not torch.equal(gathered[r], gathered[0]):
                bad = int((gathered[r] != gathered[0]).sum())
                raise SystemExit("step %d: rank %d parameters differ from rank 0 in %d checksums" % (it, r, bad))
        losses = [torch.empty_like(loss.detach().reshape(1)) for _ in range(world)]
        dist.all_gather(losses, loss.detach().reshape(1))
        if rank == 0:
            print("step %d (%s): parameters identical on %d ranks; local losses %s" % (
                it, "graph" if it >= 4 else "eager", world, [round(float(l), 5) for l in losses]), flush=True)
    if exact:
        assert len({round(float(l), 6) for l in losses}) == 1, "exact-DP Dice: every rank must report the GLOBAL loss"
    else:
        assert len({round(float(l), 7) for l in losses}) > 1, "ranks saw the same data: the check proves nothing"
    dist.barrier()
    if rank == 0:
        print("DDP CHECK OK")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
