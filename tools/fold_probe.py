"""GPU timing of fpl_wgrad_tapmajor_to_dw_batch over all conv layers of the benchmark net (one launch).  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops

DEV = "cuda:0"
L = lib.load()
ft = [16, 32, 64, 128, 256]
layers = []
for i in range(5):
    cin = 1 if i == 0 else ft[i - 1]
    if i > 0:
        layers.append((ft[i], cin, 27))
    layers.append((ft[i], ft[i], 27))
for lvl in range(4):
    layers += [(ft[lvl], 2 * ft[lvl], 27), (ft[lvl], ft[lvl], 27), (ft[lvl], ft[lvl + 1], -8)]
layers.append((2, 16, 9))
scr = [torch.randn(abs(t) * max(co, 8) * ci, device=DEV) for co, ci, t in layers]
dws = [torch.zeros(co * ci * abs(t), device=DEV) for co, ci, t in layers]
m = len(layers)
arr_s = (ctypes.c_void_p * m)(*[t.data_ptr() for t in scr])
arr_d = (ctypes.c_void_p * m)(*[t.data_ptr() for t in dws])
ints = [(ctypes.c_int * m)(*[l[k] for l in layers]) for k in (0, 1, 2)]
scout = (ctypes.c_int * m)(*[max(l[0], 8) for l in layers])
st = ops.stream_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(12):
    if it == 2:
        e0.record()
    ops.call("fpl_wgrad_tapmajor_to_dw_batch", m, arr_s, arr_d, ints[0], ints[1], ints[2], scout, st)
e1.record()
torch.cuda.synchronize()
print("fold of %d layers, %.1f M floats: %.1f us" % (m, sum(d.numel() for d in dws) / 1e6, e0.elapsed_time(e1) / 10 * 1e3))
