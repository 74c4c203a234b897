"""tests/golden/net_modes.npz: the REFERENCE network (PyMIC/pymic/net/net3d/unet2d5_dsbn.py) in its other two modes --
2.5-D (`conv_dims` containing 2: unet2d5_dsbn.py:65-73, 110-127, 159-169) and `bilinear = True` up-sampling (:170-176) --
eval logits, train-mode logits and a few gradients, to pin the oracle's restatement of those branches
(tests/test_oracle_golden.py::test_network_modes_against_reference).  Build container only (needs /root/reference):
    python -m oracle.gen_golden_modes"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from oracle.gen_golden import NET_PARAMS, _import_reference, _ref_net  # noqa: E402

SHAPE = (16, 32, 32)
MODES = {"d25": dict(conv_dims=[2, 2, 3, 3, 3], bilinear=False), "bil3d": dict(conv_dims=[3, 3, 3, 3, 3], bilinear=True),
         "bil25": dict(conv_dims=[2, 2, 3, 3, 3], bilinear=True)}
GRADS = {"d25": ["block0.conv.conv2d_1.weight", "up4.trans2d.weight", "up1.trans3d.weight", "out_conv.weight"],
         "bil3d": ["up1.conv3d.weight", "up4.conv3d.weight", "up4.conv3d.bias", "out_conv.weight"],
         "bil25": ["up4.conv2d.weight", "up3.conv2d.bias", "up2.conv3d.weight", "block1.conv.conv2d_2.weight"]}


def main():
    UNet, Dice, CE, _ = _import_reference()
    x = torch.from_numpy(synth.synth_image(1, 1, SHAPE, seed=7))
    lab = synth.synth_label(1, 2, SHAPE, seed=7)
    y = torch.from_numpy(synth.one_hot(lab, 2))
    dom = torch.ones(1, dtype=torch.long)
    out = {}
    for tag, over in MODES.items():
        params = dict(NET_PARAMS, dropout=[0.0] * 5, **over)
        net = _ref_net(UNet, params)
        net.eval()
        with torch.no_grad():
            out[tag + "_eval_logits"] = net(x, domain_label=dom).numpy()
        net = _ref_net(UNet, params)
        net.train()
        logits = net(x, domain_label=dom)
        d = {"prediction": logits, "ground_truth": y}
        loss = 0.5 * Dice({})(d) + 0.5 * CE({})(d)
        loss.backward()
        out[tag + "_train_logits"] = logits.detach().numpy()
        out[tag + "_train_loss"] = np.asarray(loss.item(), np.float64)
        named = dict(net.named_parameters())
        for k in GRADS[tag]:
            g = named[k].grad.numpy()
            out[tag + "_gradnorm_" + k] = np.asarray(np.linalg.norm(g.astype(np.float64)), np.float64)
            out[tag + "_grad_" + k] = g.reshape(-1)[:4096].copy()          # a prefix keeps the fixture small
        out[tag + "_n_with_grad"] = np.asarray(sum(p.numel() for p in net.parameters() if p.grad is not None), np.int64)
    path = os.path.join(ROOT, "tests", "golden", "net_modes.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if "logits" in k})


if __name__ == "__main__":
    main()
