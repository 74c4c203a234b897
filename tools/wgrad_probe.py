"""GPU probe for fpl_conv3d_wgrad_tc: tries the descriptor / TMEM-layout variants and prints the error of
each against torch autograd (fp32 CPU).  Development tool, not part of the product path."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fplplus_b200 import lib, ops
from tests._util import bf16_round, max_rel, randn, to_c8

DEV = "cuda:0"


def ref_wgrad(x, dy, cin, cout, kd):
    wt = torch.zeros(cout, cin, kd, 3, 3, requires_grad=True)
    F.conv3d(x, wt, None, padding=(kd // 2, 1, 1)).backward(dy)
    return wt.grad


def run(cin, cout, kd, shape, knobs):
    n, d, h, w = shape
    x = bf16_round(randn(41, n, cin, d, h, w))
    dy = bf16_round(randn(42, n, cout, d, h, w))
    ref = ref_wgrad(x, dy, cin, cout, kd)
    L = lib.load()
    for k, v in knobs.items():
        L.fpl_debug_set(k, v)
    xb, dyb = to_c8(x.to(DEV)), to_c8(dy.to(DEV))
    dw = torch.zeros_like(ref, device=DEV)
    ops.call("fpl_conv3d_wgrad_tc", ops.ptr(xb), cin // 8, 0, ops.ptr(dyb), cout // 8, 0, ops.ptr(dw), n, d, h, w, cin, cout,
             kd, ops.stream_ptr())
    torch.cuda.synchronize()
    return max_rel(dw.cpu(), ref), dw.cpu(), ref


if __name__ == "__main__":
    cases = [(16, 16, 3, (2, 3, 16, 16)), (32, 16, 3, (1, 2, 8, 8)), (16, 32, 1, (1, 2, 16, 16)),
             (64, 48, 3, (1, 2, 6, 10)), (128, 64, 3, (1, 2, 8, 8)), (256, 256, 3, (1, 2, 8, 8)),
             (16, 16, 3, (1, 4, 64, 64)), (32, 16, 3, (1, 2, 128, 128))]
    for swap in (0, 1):
        for allow64 in (0, 1):
            for quad in ((0, 1) if allow64 else (1,)):
                print("== swap_lbo_sbo=%d allow_m64=%d m64_quadrant=%d" % (swap, allow64, quad), flush=True)
                for cin, cout, kd, shape in cases:
                    try:
                        e, got, ref = run(cin, cout, kd, shape, {10: swap, 11: allow64, 12: quad})
                        print("   cin %3d cout %3d kd %d shape %-18s max_rel %.3e" % (cin, cout, kd, shape, e), flush=True)
                    except Exception as ex:
                        print("   cin %3d cout %3d kd %d shape %-18s ERROR %s" % (cin, cout, kd, shape, ex), flush=True)
                        if "CUDA error" in str(ex):
                            sys.exit(1)
