"""GPU timing of the stem conv forward / weight gradient at the benchmark shape: tensor-core kernels with the operand built
in shared memory (stem_tc.cu, fpl_debug_set 52 = 1) against patch9 + k(3,1,1) conv / wgrad (the previous path).
Graph-timed (10 launches back to back).  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from fplplus_b200 import lib, ops
from probe_util import graph_time

DEV = "cuda:0"
L = lib.load()
p = ops.ptr
n, d, h, w, cout = 4, 32, 128, 128, 16
x = torch.randn((n, 1, d, h, w), device=DEV)
wt = torch.randn(cout, 1, 3, 3, 3, device=DEV) * 0.2
b = torch.zeros(cout, device=DEV)
y = torch.empty((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
dy = torch.randn((n, d, cout // 8, h, w, 8), device=DEV).to(torch.bfloat16)
stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
dw = torch.zeros(cout, 1, 3, 3, 3, device=DEV)
st = ops.stream_ptr
for knob in (1, 0):
    L.fpl_debug_set(52, knob)
    tf = graph_time(lambda: ops.call("fpl_stem_conv_fwd", p(x), p(wt), p(b), p(y), cout // 8, 0, p(stats), n, 1, d, h, w, cout, 3, st()))
    tb = graph_time(lambda: ops.call("fpl_stem_conv_wgrad", p(x), p(dy), cout // 8, 0, p(dw), n, 1, d, h, w, cout, 3, st()))
    print("stem_tc=%d  fpl_stem_conv_fwd %.1f us   fpl_stem_conv_wgrad %.1f us" % (knob, tf, tb), flush=True)
L.fpl_debug_set(52, 1)
# the patch-tensor path of the train step before stem_tc.cu
xs = torch.empty((n, d, 4, h, w, 8), dtype=torch.bfloat16, device=DEV)
img = torch.empty(L.fpl_conv3d_dfold_image_bytes(16, 16) * 4, dtype=torch.uint8, device=DEV)
w48 = torch.zeros(cout, 48, 3, device=DEV)
ops.call("fpl_conv3d_k311_prep_weight", p(w48), 48, cout, p(img), st())
t1 = graph_time(lambda: ops.call("fpl_patch9_c8", p(x), p(xs), 1, n, d, h, w, st()))
t2 = graph_time(lambda: ops.call("fpl_conv3d_tc_k311", p(xs), 4, 0, p(img), p(b), p(y), cout // 8, 0, p(stats), n, d, h, w, 48, cout, 32, st()))
dw16 = torch.zeros(cout, 16, 3, device=DEV)
t3 = graph_time(lambda: ops.call("fpl_conv3d_wgrad_tc_k311", p(xs), 4, 0, p(dy), cout // 8, 0, p(dw16), n, d, h, w, 16, cout, st()))
print("patch9 %.1f us + k311 conv %.1f us = %.1f us forward;  k311 wgrad %.1f us" % (t1, t2, t1 + t2, t3))
