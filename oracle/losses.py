"""Oracle: pixel/image-weighted Dice + CE, entropy regulariser, hard-Dice metric
(test infrastructure).

References: DiceLoss PyMIC/pymic/loss/seg/dice.py:20-57, get_classwise_dice
loss/seg/util.py:85-107, CrossEntropyLoss loss/seg/ce.py:23-44, CombinedLoss
loss/seg/combined.py:34-39, entropy term net_run_dsbn/agent_seg.py:467,
training-time hard Dice agent_seg.py:472-476.

Two forms: a torch form (autograd supplies gradients) and a float64 numpy
closed form of value + d/dlogits used to check the fused
CUDA kernel to 1e-5.
"""
import numpy as np
import torch


def _flat(x):
    """[N,C,D,H,W] -> [V,C] (util.py:36-50)."""
    c = x.shape[1]
    return x.permute(0, 2, 3, 4, 1).reshape(-1, c)


def dice_loss(logits, soft_y, pixel_weight=None, softmax=True):
    p = torch.softmax(logits, dim=1) if softmax else logits
    p, y = _flat(p), _flat(soft_y)
    if pixel_weight is None:
        yv, pv, it = y.sum(0), p.sum(0), (y * p).sum(0)
    else:
        w = _flat(pixel_weight)
        yv, pv, it = (y * w).sum(0), (p * w).sum(0), (y * p * w).sum(0)
    dice = (2.0 * it + 1e-5) / (yv + pv + 1e-5)
    return 1.0 - dice.mean()


def ce_loss(logits, soft_y, pixel_weight=None, softmax=True):
    p = torch.softmax(logits, dim=1) if softmax else logits
    p, y = _flat(p), _flat(soft_y)
    p = p * 0.999 + 5e-4
    ce = -(y * torch.log(p)).sum(1)
    if pixel_weight is None:
        return ce.mean()
    w = _flat(pixel_weight).squeeze()
    return (w * ce).sum() / (w.sum() + 1e-5)


def combined_loss(logits, soft_y, pixel_weight=None, w_dice=1.0, w_ce=0.0):
    """loss_type = [DiceLoss, CrossEntropyLoss], loss_weight = [w_dice, w_ce];
    loss_type = DiceLoss alone is (1, 0)."""
    val = 0.0
    if w_dice != 0.0:
        val = val + w_dice * dice_loss(logits, soft_y, pixel_weight)
    if w_ce != 0.0:
        val = val + w_ce * ce_loss(logits, soft_y, pixel_weight)
    return val


def entropy_bits(logits):
    """-(p*log2(p+1e-10)).sum()/(N*D*H*W) -- agent_seg.py:466-467 (the names
    D,B,C,W,H there are mislabelled; the divisor is all dims but the class one)."""
    p = logits.softmax(1)
    n, c, d, h, w = logits.shape
    return -(p * torch.log2(p + 1e-10)).sum() / (n * d * h * w)


def hard_dice(logits, soft_y):
    """argmax -> one-hot -> classwise Dice (agent_seg.py:472-476)."""
    c = logits.shape[1]
    am = torch.argmax(logits, dim=1)
    oh = torch.nn.functional.one_hot(am, c).permute(0, 4, 1, 2, 3).to(soft_y.dtype)
    o, y = _flat(oh), _flat(soft_y)
    return (2.0 * (o * y).sum(0) + 1e-5) / (y.sum(0) + o.sum(0) + 1e-5)


def dice_ce_closed_form(logits, soft_y, pixel_weight=None, w_dice=1.0, w_ce=0.0):
    """float64 numpy value and gradient wrt logits of w_dice*Dice + w_ce*CE.
    Returns (loss, dlogits[N,C,D,H,W], sums) with sums = (I_c, Y_c, P_c, sum_w, sum_w_ce)."""
    z = np.asarray(logits, np.float64)
    y = np.asarray(soft_y, np.float64)
    n, c = z.shape[:2]
    w = np.ones((n, 1) + z.shape[2:]) if pixel_weight is None else np.asarray(pixel_weight, np.float64)
    zs = z - z.max(axis=1, keepdims=True)
    e = np.exp(zs)
    p = e / e.sum(axis=1, keepdims=True)
    ax = (0, 2, 3, 4)
    I, Y, P = (w * y * p).sum(ax), (w * y).sum(ax), (w * p).sum(ax)
    eps = 1e-5
    den = Y + P + eps
    dice = (2 * I + eps) / den
    ldice = 1.0 - dice.mean()
    q = 0.999 * p + 5e-4
    ce_v = -(y * np.log(q)).sum(axis=1, keepdims=True)
    if pixel_weight is None:
        cden = float(z.size // c)
    else:
        cden = w.sum() + eps
    lce = (w * ce_v).sum() / cden
    loss = w_dice * ldice + w_ce * lce
    sh = (1, c, 1, 1, 1)
    g = np.zeros_like(z)
    if w_dice != 0.0:
        g += w_dice * (-(1.0 / c) * w * (2 * y * den.reshape(sh) - (2 * I + eps).reshape(sh)) / (den.reshape(sh) ** 2))
    if w_ce != 0.0:
        g += w_ce * (-0.999 * w * y / q / cden)
    dz = p * (g - (g * p).sum(axis=1, keepdims=True))
    return loss, dz, (I, Y, P, w.sum(), (w * ce_v).sum())
