"""GPU probe for the tensor-core transposed-conv kernels against torch (fp32 CPU).  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from fplplus_b200 import lib, ops
from tests._util import bf16_round, from_c8, max_rel, randn, to_c8

DEV = "cuda:0"
L = lib.load()
for cin, cout, kd2, shape in [(32, 16, 2, (2, 2, 16, 8)), (64, 32, 1, (1, 3, 20, 12)), (256, 128, 2, (1, 1, 4, 6)),
                              (128, 64, 2, (1, 2, 8, 8)), (32, 16, 2, (1, 8, 64, 64))]:
    n, d, h, w = shape
    x = bf16_round(randn(71, n, cin, d, h, w)).requires_grad_(True)
    wt = bf16_round(randn(72, cin, cout, kd2, 2, 2, scale=0.2)).requires_grad_(True)
    b = randn(73, cout, scale=0.1)
    ref = F.conv_transpose3d(x, wt, b, stride=(kd2, 2, 2))
    g = bf16_round(randn(74, *ref.shape))
    ref.backward(g)
    xb = to_c8(x.detach().to(DEV))
    wd, bd = wt.detach().to(DEV), b.to(DEV)
    nbytes = L.fpl_convt_weight_image_bytes(cin, cout, kd2)
    img = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=DEV)
    img_t = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=DEV)
    st = ops.stream_ptr()
    ops.call("fpl_convt_prep_weight", ops.ptr(wd), cin, cout, kd2, 0, ops.ptr(img), st)
    ops.call("fpl_convt_prep_weight", ops.ptr(wd), cin, cout, kd2, 1, ops.ptr(img_t), st)
    do, ho, wo = d * kd2, 2 * h, 2 * w
    cat = torch.zeros((n, do, 2 * cout // 8, ho, wo, 8), dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_convt_k2s2_fwd_tc", ops.ptr(xb), cin // 8, 0, ops.ptr(img), ops.ptr(bd), ops.ptr(cat), 2 * cout // 8, cout // 8,
             n, d, h, w, cin, cout, kd2, st)
    torch.cuda.synchronize()
    out = from_c8(cat).cpu()
    e_f = max_rel(out[:, cout:], ref.detach())
    untouched = bool(torch.all(out[:, :cout] == 0))
    gcat = to_c8(torch.cat([torch.zeros_like(g), g], 1).to(DEV))
    dx = torch.zeros((n, d, cin // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_convt_k2s2_dgrad_tc", ops.ptr(gcat), 2 * cout // 8, cout // 8, ops.ptr(img_t), ops.ptr(dx), cin // 8, 0,
             n, d, h, w, cin, cout, kd2, st)
    dw = torch.zeros(cin, cout, kd2, 2, 2, device=DEV)
    ops.call("fpl_convt_k2s2_wgrad_tc", ops.ptr(xb), cin // 8, 0, ops.ptr(gcat), 2 * cout // 8, cout // 8, ops.ptr(dw),
             n, d, h, w, cin, cout, kd2, st)
    torch.cuda.synchronize()
    e_d = max_rel(from_c8(dx).cpu(), x.grad)
    e_w = max_rel(dw.cpu(), wt.grad)
    print("cin %3d cout %3d kd2 %d %-16s fwd %.2e (other half untouched %s) dgrad %.2e wgrad %.2e" %
          (cin, cout, kd2, shape, e_f, untouched, e_d, e_w), flush=True)
