"""GPU probe: tensor-core conv vs direct conv on a few shapes, for both LBO/SBO conventions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fplplus_b200 import ops, lib as L
from tests._util import randn, bf16_round, to_c8, from_c8, max_rel

DEV = "cuda:0"
lib = L.load()
print("sm100:", lib.fpl_device_is_sm100(), torch.cuda.get_device_name(0))
for swap in (0, 1):
    lib.fpl_debug_set(0, swap)
    for (cin, cout, kd, shape) in [(16, 16, 3, (1, 2, 16, 8)), (16, 16, 1, (1, 1, 16, 8)), (32, 32, 3, (1, 3, 24, 24)), (64, 64, 3, (2, 3, 16, 8))]:
        n, d, h, w = shape
        x = bf16_round(randn(11, n, cin, d, h, w)); wt = bf16_round(randn(12, cout, cin, kd, 3, 3, scale=0.1))
        xb = to_c8(x.to(DEV)); wd = wt.to(DEV)
        img = torch.empty(cin * cout * kd * 9, dtype=torch.bfloat16, device=DEV)
        ops.call("fpl_conv3d_prep_weight", ops.ptr(wd), cin, cout, kd, 0, ops.ptr(img), ops.stream_ptr())
        y = torch.zeros((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
        y2 = torch.zeros_like(y)
        try:
            ops.call("fpl_conv3d_tc", ops.ptr(xb), cin // 8, 0, ops.ptr(img), None, ops.ptr(y), cout // 8, 0, None, n, d, h, w, cin, cout, kd, ops.stream_ptr())
            torch.cuda.synchronize()
        except Exception as e:
            print("swap", swap, (cin, cout, kd, shape), "ERROR", e); continue
        ops.call("fpl_conv3d_direct", ops.ptr(xb), cin // 8, 0, ops.ptr(wd), None, ops.ptr(y2), cout // 8, 0, None, n, d, h, w, cin, cout, kd, 0, 1, ops.stream_ptr())
        torch.cuda.synchronize()
        a, b = from_c8(y).cpu(), from_c8(y2).cpu()
        print("swap", swap, (cin, cout, kd, shape), "max_rel tc vs direct:", max_rel(a, b), "nonzero frac", float((a != 0).float().mean()))
lib.fpl_debug_set(0, 0)
