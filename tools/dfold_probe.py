"""GPU probe for the depth-folded conv kernel against torch (fp32 CPU) and timing against fpl_conv3d_tc."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fplplus_b200 import lib, ops
from tests._util import bf16_round, from_c8, max_rel, randn, to_c8

DEV = "cuda:0"
L = lib.load()


def run(cin, cout, shape, transpose, timing=False):
    n, d, h, w = shape
    x = bf16_round(randn(11, n, cin, d, h, w))
    wt = bf16_round(randn(12, cout, cin, 3, 3, 3, scale=0.1))
    b = randn(13, cout, scale=0.1)
    st = ops.stream_ptr()
    if not transpose:
        ref = F.conv3d(x, wt, b, padding=1)
        xin, ci, co, bias = x, cin, cout, b.to(DEV)
    else:   # dgrad of conv(cin -> cout): input dy has cout channels, output cin
        dy = bf16_round(randn(14, n, cout, d, h, w))
        xx = torch.zeros(n, cin, d, h, w, requires_grad=True)
        F.conv3d(xx, wt, None, padding=1).backward(dy)
        ref, xin, ci, co, bias = xx.grad, dy, cout, cin, None
    xb = to_c8(xin.to(DEV))
    wd = wt.to(DEV)
    nbytes = L.fpl_conv3d_dfold_image_bytes(ci, co)
    assert nbytes > 0, (ci, co)
    img = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_conv3d_dfold_prep_weight", ops.ptr(wd), cin, cout, 1 if transpose else 0, ops.ptr(img), st)
    y = torch.zeros((n, d, co // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2 * co, dtype=torch.float64, device=DEV)
    ops.call("fpl_conv3d_tc_dfold", ops.ptr(xb), ci // 8, 0, ops.ptr(img), ops.ptr(bias), ops.ptr(y), co // 8, 0, ops.ptr(stats),
             n, d, h, w, ci, co, st)
    torch.cuda.synchronize()
    out = from_c8(y).cpu()
    e = max_rel(out, ref.detach())
    es = float((stats.cpu()[:co] - ref.detach().double().sum((0, 2, 3, 4))).abs().max() / (ref.detach().double().sum((0, 2, 3, 4)).abs().max() + 1e-9))
    msg = "cin %3d cout %3d %-18s %s max_rel %.2e stats %.1e" % (cin, cout, shape, "dgrad" if transpose else "fwd  ", e, es)
    if timing:
        img2 = torch.empty(L.fpl_conv3d_weight_image_bytes(ci, co, 3) // 2, dtype=torch.bfloat16, device=DEV)
        ops.call("fpl_conv3d_prep_weight", ops.ptr(wd), cin, cout, 3, 1 if transpose else 0, ops.ptr(img2), st)
        res = []
        for which in (0, 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(6):
                if it == 1:
                    e0.record()
                if which == 0:
                    ops.call("fpl_conv3d_tc_dfold", ops.ptr(xb), ci // 8, 0, ops.ptr(img), ops.ptr(bias), ops.ptr(y), co // 8, 0,
                             ops.ptr(stats), n, d, h, w, ci, co, st)
                else:
                    ops.call("fpl_conv3d_tc", ops.ptr(xb), ci // 8, 0, ops.ptr(img2), ops.ptr(bias), ops.ptr(y), co // 8, 0,
                             ops.ptr(stats), n, d, h, w, ci, co, 3, st)
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 5 * 1e3)
        msg += "   dfold %.1f us vs conv3d_tc %.1f us" % tuple(res)
    print(msg, flush=True)


for cin, cout, shape in [(16, 16, (2, 4, 32, 16)), (32, 16, (1, 3, 20, 12)), (16, 32, (1, 17, 16, 8)), (32, 32, (1, 5, 24, 24)),
                         (64, 32, (1, 2, 16, 16)), (16, 16, (1, 20, 16, 16))]:
    run(cin, cout, shape, False)
    run(cin, cout, shape, True)
for cin, cout, shape in [(16, 16, (4, 32, 128, 128)), (32, 16, (4, 32, 128, 128)), (32, 32, (4, 16, 64, 64)), (64, 32, (4, 16, 64, 64))]:
    run(cin, cout, shape, False, timing=True)
run(32, 16, (4, 32, 128, 128), True, timing=True)
