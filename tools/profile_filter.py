"""One launch of every filter / stitching / loss kernel at the configs[1] / configs[2] sizes between
cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --set full --clock-control none -o gpurun_out/r02_filter python tools/profile_filter.py
Numbers printed under a profiler are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from fplplus_b200 import fpl
from fplplus_b200.ops import call, ptr, stream_ptr


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    d, h, w = bench.VOLUME
    C, K = 2, 6
    g = torch.Generator().manual_seed(5)
    vols = [torch.randn((1, C, d, h, w), generator=g).to(dev) for _ in range(K + 2)]
    n, sp = bench.BATCH, bench.PATCH[0] * bench.PATCH[1] * bench.PATCH[2]
    z = torch.randn((n, C) + bench.PATCH, generator=g).to(dev)
    lab = torch.randint(0, C, (n,) + bench.PATCH, generator=g, dtype=torch.uint8).to(dev)
    code = torch.randint(1, 3, (n, 1) + bench.PATCH, generator=g, dtype=torch.uint8).to(dev)
    onehot = torch.nn.functional.one_hot(lab.long(), C).permute(0, 4, 1, 2, 3).float().contiguous()
    pw = code.float() * 0.5
    iw = torch.rand(n, generator=g).to(dev)
    sums = torch.zeros(6 * C + 3, dtype=torch.float64, device=dev)
    dz = torch.empty_like(z)
    gs = torch.ones((), device=dev)
    acc, cnt = torch.zeros((1, C, d, h, w), device=dev), torch.zeros((1, C, d, h, w), device=dev)
    patch = torch.randn((1, C, 32, 128, 128), generator=g).to(dev)

    def run():
        fpl.mc_uncertainty(vols[:K])
        fpl.agreement_weight(vols[0], vols[1])
        fpl.pseudo_label(vols[2])
        call("fpl_window_accumulate", ptr(patch), ptr(acc), ptr(cnt), 1, C, d, h, w, 0, 0, 0, 32, 128, 128, 0, 1, 1.0, stream_ptr())
        call("fpl_window_normalize", ptr(acc), ptr(cnt), 1.0, acc.numel(), stream_ptr())
        call("fpl_dice_ce_reduce_ex", ptr(z), None, ptr(lab), None, ptr(code), ptr(iw), ptr(sums), n, C, sp, 0, 0, stream_ptr())
        call("fpl_dice_ce_grad_ex", ptr(z), None, ptr(lab), None, ptr(code), ptr(iw), ptr(sums), 0.5, 0.5, 0.0, 1.0, ptr(gs), None,
             ptr(dz), n, C, sp, 0, 0, stream_ptr())
        call("fpl_dice_ce_reduce_ex", ptr(z), ptr(onehot), None, ptr(pw), None, None, ptr(sums), n, C, sp, 0, 0, stream_ptr())
        call("fpl_dice_ce_grad_ex", ptr(z), ptr(onehot), None, ptr(pw), None, None, ptr(sums), 0.5, 0.5, 0.0, 1.0, ptr(gs), None,
             ptr(dz), n, C, sp, 0, 0, stream_ptr())
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
