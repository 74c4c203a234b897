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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    agent = bench.build_agent("train", world)
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
    assert len({round(float(l), 7) for l in losses}) > 1, "ranks saw the same data: the check proves nothing"
    dist.barrier()
    if rank == 0:
        print("DDP CHECK OK")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
