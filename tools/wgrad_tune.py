"""GPU timing of fpl_conv3d_wgrad_tc on the network's layer shapes under the tuning knobs.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops

DEV = "cuda:0"
L = lib.load()
SHAPES = [(16, 8, 1, (4, 32, 128, 128)), (16, 16, 3, (4, 32, 128, 128)), (32, 16, 3, (4, 32, 128, 128)), (16, 16, 1, (4, 32, 128, 128)),
          (16, 32, 3, (4, 16, 64, 64)), (32, 32, 3, (4, 16, 64, 64)), (64, 32, 3, (4, 16, 64, 64)),
          (32, 64, 3, (4, 8, 32, 32)), (64, 64, 3, (4, 8, 32, 32)), (128, 64, 3, (4, 8, 32, 32)),
          (64, 128, 3, (4, 4, 16, 16)), (128, 128, 3, (4, 4, 16, 16)), (256, 128, 3, (4, 4, 16, 16)),
          (128, 256, 3, (4, 2, 8, 8)), (256, 256, 3, (4, 2, 8, 8))]


def t(cin, cout, kd, shape, knobs):
    n, d, h, w = shape
    for k, v in knobs.items():
        L.fpl_debug_set(k, v)
    x = torch.randn((n, d, cin // 8, h, w, 8), device=DEV).to(torch.bfloat16)
    dy = torch.randn((n, d, max(cout // 8, 2), h, w, 8), device=DEV).to(torch.bfloat16)
    dw = torch.zeros(cout, cin, kd, 3, 3, device=DEV)
    st = ops.stream_ptr()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(7):
        if it == 2:
            e0.record()
        ops.call(FN, ops.ptr(x), cin // 8, 0, ops.ptr(dy), max(cout // 8, 2), 0, ops.ptr(dw), n, d, h, w, cin, cout, kd, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3


variants = [("tiles2", {16: 2, 17: 0}), ("t2-noepi", {16: 2, 17: 1}), ("tiles4", {16: 4, 17: 0}), ("tiles8", {16: 8, 17: 0}),
            ("tiles32", {16: 32, 17: 0}), ("t32-noepi", {16: 32, 17: 1})]
FN = os.environ.get("FN", "fpl_conv3d_wgrad_tc_tapmajor")
if os.environ.get("DEEP_ONLY"):
    SHAPES = [s_ for s_ in SHAPES if s_[3][2] <= 32]
print("%-28s" % "shape" + "".join("%12s" % v[0] for v in variants))
for cin, cout, kd, shape in SHAPES:
    row = []
    for name, knobs in variants:
        try:
            row.append("%10.1fus" % t(cin, cout, kd, shape, knobs))
        except Exception as ex:
            row.append("%12s" % "n/a")
    print("%-28s" % ("%d->%d k%d %s" % (cin, cout, kd, "x".join(map(str, shape[1:])))) + "".join(row), flush=True)
L.fpl_debug_set(16, 2)
L.fpl_debug_set(17, 0)
