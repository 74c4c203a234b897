"""GPU probe: per-unit comparison of the CUDA forward (conv outputs y, activations a) with the
bf16-emulating oracle, to localise semantic mismatches.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from oracle import synth, unet_dsbn
from oracle.gen_golden import NET_PARAMS, SHAPE
from fplplus_b200 import ops
from fplplus_b200.net import UNet2D5_dsbn, _Workspace

DEV = "cuda:0"
params = dict(NET_PARAMS, dropout=[0.0] * 5)
sd = synth.synth_state_dict()
net = UNet2D5_dsbn(dict(params))
net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
net = net.to(DEV).train()
x = torch.from_numpy(synth.synth_image(2, 1, SHAPE, seed=1))
ws = _Workspace(torch.device(DEV))
with torch.no_grad():
    logits, rec = net._run_forward(x.to(DEV), 1, ws)
torch.cuda.synchronize()

# emulation with taps
st = unet_dsbn.to_torch_state(sd)
taps = {}
orig_bn = unet_dsbn._bn


def bn_tap(state, prefix, xx, domain, training, bf16=False):
    taps["Y:" + prefix] = xx.detach().clone()
    return orig_bn(state, prefix, xx, domain, training, bf16)


unet_dsbn._bn = bn_tap
with torch.no_grad():
    emu = unet_dsbn.forward(st, x, 1, params, bn_training=True, bf16=True)


def rl(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


names = [("block%d.conv" % i) for i in range(5)] + [("up%d.conv" % k) for k in (1, 2, 3, 4)]
for nm in names:
    for k in (1, 2):
        ours_y = ops.c8_to_ncdhw(rec["%s#%d" % (nm, k)]["y"]).cpu()
        ref_y = taps["Y:%s.bn3d%d" % (nm, k)]
        r = rec["%s#%d" % (nm, k)]
        print("%-14s#%d  y rel_l2 %.5f  (vs rounded emu %.5f)  max|y| %.2f" % (
            nm, k, rl(ours_y, ref_y), rl(ours_y, ref_y.to(torch.bfloat16).float()), float(ref_y.abs().max())))
print("logits rel_l2", rl(logits.cpu(), emu))
