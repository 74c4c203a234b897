"""The bar on the same box: the FPL+ DSBN 3-D U-Net train step written with stock ``torch.nn`` modules, i.e. what the
reference's own network (PyMIC/pymic/net/net3d/unet2d5_dsbn.py:48-309 with conv_dims all 3, bilinear=False) executes
on a B200 through cuDNN / cuBLAS / ATen kernels.  BASELINE ONLY: bench.py times it beside the hand-written path
(`torch_cuda_baseline` in the JSON line); nothing in fplplus_b200 imports it and it is not a fallback.

Same architecture and step semantics as the product arm: two domain-specific BatchNorm3d per conv, PReLU, Dropout,
MaxPool3d, ConvTranspose3d k2s2, (1,3,3) head; one `training_all` step = zero_grad, forward + 0.5 Dice + 0.5 CE
(pixel-weighted on the target batch) for both domains, backward, Adam(weight_decay).  Modes: fp32 NCDHW (what the
reference .cfg runs: tensor_type = float) and bf16 autocast + channels_last_3d (the fastest stock configuration)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _DSBN(nn.Module):
    def __init__(self, c, nd):
        super().__init__()
        self.bns = nn.ModuleList([nn.BatchNorm3d(c) for _ in range(nd)])

    def forward(self, x, d):
        return self.bns[d](x)


class _Block(nn.Module):
    def __init__(self, ci, co, nd, p):
        super().__init__()
        self.c1, self.c2 = nn.Conv3d(ci, co, 3, padding=1), nn.Conv3d(co, co, 3, padding=1)
        self.b1, self.b2 = _DSBN(co, nd), _DSBN(co, nd)
        self.r1, self.r2 = nn.PReLU(), nn.PReLU()
        self.drop = nn.Dropout(p)

    def forward(self, x, d):
        x = self.drop(self.r1(self.b1(self.c1(x), d)))
        return self.r2(self.b2(self.c2(x), d))


class TorchUNetDSBN(nn.Module):
    def __init__(self, in_chns=1, ft=(16, 32, 64, 128, 256), n_class=2, nd=2, dropout=(0.0, 0.0, 0.3, 0.4, 0.5)):
        super().__init__()
        ch = [in_chns] + list(ft)
        self.down = nn.ModuleList([_Block(ch[i], ch[i + 1], nd, dropout[i]) for i in range(5)])
        self.trans = nn.ModuleList([nn.ConvTranspose3d(ft[4 - k], ft[3 - k], 2, stride=2) for k in range(4)])
        self.up = nn.ModuleList([_Block(2 * ft[3 - k], ft[3 - k], nd, dropout[3 - k]) for k in range(4)])
        self.head = nn.Conv3d(ft[0], n_class, (1, 3, 3), padding=(0, 1, 1))

    def forward(self, x, d):
        skips = []
        for i, blk in enumerate(self.down):
            x = blk(x, d)
            if i < 4:
                skips.append(x)
                x = F.max_pool3d(x, 2)
        for k in range(4):
            x = self.up[k](torch.cat([skips[3 - k], self.trans[k](x)], 1), d)
        return self.head(x)


def dice_ce(logits, y, w, w_dice=0.5, w_ce=0.5):
    """loss/seg/dice.py:20-57 + ce.py:23-44 + combined.py:34-39 in stock torch ops (fp32)."""
    p = torch.softmax(logits.float(), 1)
    c = p.shape[1]
    pf, yf = p.permute(0, 2, 3, 4, 1).reshape(-1, c), y.permute(0, 2, 3, 4, 1).reshape(-1, c)
    if w is None:
        yv, pv, it = yf.sum(0), pf.sum(0), (yf * pf).sum(0)
    else:
        wf = w.reshape(-1, 1)
        yv, pv, it = (yf * wf).sum(0), (pf * wf).sum(0), (yf * pf * wf).sum(0)
    dice = 1.0 - ((2.0 * it + 1e-5) / (yv + pv + 1e-5)).mean()
    ce = -(yf * torch.log(pf * 0.999 + 5e-4)).sum(1)
    ce = ce.mean() if w is None else (w.reshape(-1) * ce).sum() / (w.sum() + 1e-5)
    return w_dice * dice + w_ce * ce


class TorchTrainer(object):
    def __init__(self, device, mode="fp32", lr=1e-4, weight_decay=1e-5):
        assert mode in ("fp32", "bf16_channels_last")
        torch.backends.cudnn.benchmark = True
        self.mode, self.device = mode, device
        self.net = TorchUNetDSBN().to(device).train()
        if mode == "bf16_channels_last":
            self.net = self.net.to(memory_format=torch.channels_last_3d)
        self.opt = torch.optim.Adam(self.net.parameters(), lr, weight_decay=weight_decay, fused=True)

    def step(self, batches):
        """batches: [(x, onehot, weight or None)] per domain, CUDA fp32 tensors."""
        self.opt.zero_grad(set_to_none=True)
        total = None
        for d, (x, y, w) in enumerate(batches):
            if self.mode == "bf16_channels_last":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    z = self.net(x.contiguous(memory_format=torch.channels_last_3d), d)
            else:
                z = self.net(x, d)
            l = dice_ce(z, y, w)
            total = l if total is None else total + l
        loss = total / len(batches)
        loss.backward()
        self.opt.step()
        return loss
