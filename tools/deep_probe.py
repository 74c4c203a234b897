"""GPU probe for the deep-level conv3d_tc launches (levels 2-4 of the bench network), with and without cluster split-K.
Under ncu:  ncu --set full --clock-control none --cache-control none --import-source on -k regex:conv3d_tc_kernel \
            -o gpurun_out/deep python tools/deep_probe.py once
Plain: graph-timed launches for ksplit on / off."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops
from tests._util import bf16_round, randn, to_c8

DEV = "cuda:0"
L = lib.load()
P = ops.ptr
CASES = [(64, 64, (4, 8, 32, 32)), (128, 64, (4, 8, 32, 32)), (64, 128, (4, 4, 16, 16)), (128, 128, (4, 4, 16, 16)),
         (256, 128, (4, 4, 16, 16)), (128, 256, (4, 2, 8, 8)), (256, 256, (4, 2, 8, 8)), (256, 256, (1, 2, 8, 8))]


def setup(cin, cout, shape):
    n, d, h, w = shape
    x = to_c8(bf16_round(randn(1, n, cin, d, h, w)).to(DEV))
    wt = bf16_round(randn(2, cout, cin, 3, 3, 3, scale=0.1)).to(DEV)
    y = torch.zeros((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    img = torch.empty(L.fpl_conv3d_weight_image_bytes(cin, cout, 3) // 2, dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_conv3d_prep_weight", P(wt), cin, cout, 3, 0, P(img), ops.stream_ptr())
    return lambda: ops.call("fpl_conv3d_tc", P(x), cin // 8, 0, P(img), None, P(y), cout // 8, 0, P(stats), n, d, h, w, cin, cout, 3,
                            ops.stream_ptr()), (x, wt, y, stats, img)


def graph_time(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters * 1e3)
    return sorted(ts)[1]


once = len(sys.argv) > 1 and sys.argv[1] == "once"
for cin, cout, shape in CASES:
    fn, keep = setup(cin, cout, shape)
    res = []
    for allow in (1, 0):
        L.fpl_debug_set(2, allow)
        if once:
            fn()
            fn()
            torch.cuda.synchronize()
        else:
            res.append(graph_time(fn))
    L.fpl_debug_set(2, 1)
    if not once:
        print("%3d->%3d %-16s split-K %6.1f us | single CTA per tile %6.1f us" % (cin, cout, shape, res[0], res[1]), flush=True)
