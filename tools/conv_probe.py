"""GPU timing of fpl_conv3d_tc on the deep-level shapes (levels 2-4, batch 4 and batch 1) with the N split of the
staged weight slices on / off (fpl_debug_set key 1).  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops

DEV = "cuda:0"
L = lib.load()
SHAPES = [(32, 64, (4, 8, 32, 32)), (64, 64, (4, 8, 32, 32)), (128, 64, (4, 8, 32, 32)), (64, 128, (4, 4, 16, 16)),
          (128, 128, (4, 4, 16, 16)), (256, 128, (4, 4, 16, 16)), (128, 256, (4, 2, 8, 8)), (256, 256, (4, 2, 8, 8)),
          (128, 128, (1, 4, 16, 16)), (256, 256, (1, 2, 8, 8)), (64, 64, (1, 8, 32, 32))]


def t(cin, cout, shape, nsub):
    n, d, h, w = shape
    L.fpl_debug_set(1, nsub)
    x = torch.randn((n, d, cin // 8, h, w, 8), device=DEV).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 3, 3, 3, device=DEV) * 0.05
    img = torch.empty(cin * cout * 27, dtype=torch.bfloat16, device=DEV)
    st = ops.stream_ptr()
    ops.call("fpl_conv3d_prep_weight", ops.ptr(wt), cin, cout, 3, 0, ops.ptr(img), st)
    y = torch.empty((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    bias = torch.zeros(cout, device=DEV)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    for _ in range(3):
        ops.call("fpl_conv3d_tc", ops.ptr(x), cin // 8, 0, ops.ptr(img), ops.ptr(bias), ops.ptr(y), cout // 8, 0, ops.ptr(stats),
                 n, d, h, w, cin, cout, 3, st)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            sp = ops.stream_ptr()
            for _ in range(20):
                ops.call("fpl_conv3d_tc", ops.ptr(x), cin // 8, 0, ops.ptr(img), ops.ptr(bias), ops.ptr(y), cout // 8, 0,
                         ops.ptr(stats), n, d, h, w, cin, cout, 3, sp)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3, y.float().clone()


print("%-30s %10s %10s" % ("shape", "nsub off", "nsub on"))
for cin, cout, shape in SHAPES:
    a, ya = t(cin, cout, shape, 0)
    b, yb = t(cin, cout, shape, 1)
    print("%-30s %8.1fus %8.1fus   max|diff| %.3g" % ("%d->%d %s" % (cin, cout, "x".join(map(str, shape))), a, b,
                                                     float((ya - yb).abs().max())), flush=True)
