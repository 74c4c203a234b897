"""Drop-in ``loss_dict`` entries: pixel/image-weighted Dice, cross entropy and their weighted sum,
computed by one fused reduction kernel + one gradient kernel (csrc/loss.cu).

API mirrors PyMIC (PyMIC/pymic/loss/seg/abstract.py:16-21, dice.py:20-57, ce.py:23-44,
combined.py:21-39): ``cls(params)`` reads ``loss_softmax`` (default True); ``forward(dict)`` takes
``prediction`` (tensor or list/tuple -> [0]), ``ground_truth`` [N,C,D,H,W] float one-hot/soft,
optional ``pixel_weight`` [N,1,D,H,W]; ``image_weight`` is accepted and ignored exactly as the
reference's DiceLoss/CrossEntropyLoss ignore it (image weights enter through
NiftyDataset.set_weight_, io/nifty_dataset.py:165-168).  Returns a 0-dim tensor attached to autograd.
"""
import torch
import torch.nn as nn

from .ops import call, ptr, stream_ptr


def _prep(loss_input_dict):
    """Returns (prediction, truth, weight): ``truth`` is fp32 one-hot/soft [N,C,D,H,W] (the PyMIC layout) or a uint8
    label map [N,D,H,W] / [N,1,D,H,W] (device data path: the one-hot is built inside the kernel); ``weight`` is None,
    an fp32 map [N,1,D,H,W], or a (uint8 agreement code [N,1,D,H,W], image weight [N] or None) pair."""
    predict = loss_input_dict['prediction']
    soft_y = loss_input_dict['ground_truth']
    pix_w = loss_input_dict.get('pixel_weight', None)
    if isinstance(predict, (list, tuple)):
        predict = predict[0]
    if not predict.is_cuda:
        raise RuntimeError("fplplus_b200 losses run on CUDA only")
    if predict.dim() != 5:
        raise ValueError("{0:}D tensor not supported".format(predict.dim()))
    n_vox = predict.numel() // predict.shape[1]
    if soft_y.dtype == torch.uint8:
        soft_y = soft_y.to(device=predict.device).contiguous()
        if soft_y.numel() != n_vox:
            raise ValueError("uint8 ground_truth must be a label map [N,D,H,W] matching prediction %s" % (tuple(predict.shape),))
    else:
        soft_y = soft_y.to(device=predict.device, dtype=torch.float32).contiguous()
        if soft_y.shape != predict.shape:
            raise ValueError("prediction %s and ground_truth %s differ in shape" % (tuple(predict.shape), tuple(soft_y.shape)))
    if pix_w is not None:
        if pix_w.dtype == torch.uint8:
            code = pix_w.to(device=predict.device).contiguous()
            if code.numel() != n_vox:
                raise ValueError("pixel_weight must be [N,1,D,H,W]")
            iw = loss_input_dict.get('image_weight', None) if loss_input_dict.get('fold_image_weight', False) else None
            if iw is not None:
                iw = torch.as_tensor(iw).to(device=predict.device, dtype=torch.float32).reshape(-1).contiguous()
                if iw.numel() != predict.shape[0]:
                    raise ValueError("image_weight must hold one value per sample")
            pix_w = (code, iw)
        else:
            pix_w = pix_w.to(device=predict.device, dtype=torch.float32).contiguous()
            if pix_w.numel() != n_vox:
                raise ValueError("pixel_weight must be [N,1,D,H,W]")
    return predict, soft_y, pix_w


def n_sums(c):
    return 6 * c + 3


def _src_args(truth, weight):
    """(soft_y, label, weight, weight_code, image_weight) pointers of fpl_dice_ce_*_ex."""
    soft_y, label = (None, truth) if truth.dtype == torch.uint8 else (truth, None)
    w, code, iw = None, None, None
    if isinstance(weight, tuple):
        code, iw = weight
    else:
        w = weight
    return ptr(soft_y), ptr(label), ptr(w), ptr(code), ptr(iw)


class _DiceCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, truth, weight, w_dice, w_ce, w_ent, prob_input, holder):
        logits = logits.float().contiguous()
        n, c = logits.shape[:2]
        spatial = logits.numel() // (n * c)
        sums = torch.zeros(n_sums(c), dtype=torch.float64, device=logits.device)
        st = stream_ptr()
        src = _src_args(truth, weight)
        call("fpl_dice_ce_reduce_ex", ptr(logits), *src, ptr(sums), n, c, spatial, 1 if w_ent != 0.0 else 0, prob_input, st)
        # exact data-parallel mode: the (6C+3) partial sums of all ranks are summed between the two passes,
        # so Dice / CE are those of the GLOBAL batch as under nn.DataParallel (agent_seg.py:695, dice.py:29-35); the
        # gradient is scaled by the world size because the parameter gradients are AVERAGED over ranks afterwards
        world = 1
        red = holder.get("sums_allreduce") if holder is not None else None
        if red is not None:
            world = int(red(sums))
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        hard = torch.empty(c, dtype=torch.float64, device=logits.device)
        call("fpl_dice_ce_loss_ex", ptr(logits), *src, ptr(sums), w_dice, w_ce, w_ent, ptr(loss), ptr(hard), n, c,
             spatial, prob_input, n * world if world > 1 else 0, st)
        ctx.saved = (logits, truth, weight, sums)
        ctx.w = (w_dice, w_ce, w_ent, prob_input, world)
        if holder is not None:
            holder["sums"] = sums
            holder["hard_dice"] = hard
            holder["voxels"] = n * spatial
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        logits, truth, weight, sums = ctx.saved
        n, c = logits.shape[:2]
        spatial = logits.numel() // (n * c)
        dlogits = torch.empty_like(logits)
        g = grad_out.float().contiguous()
        world = ctx.w[4]
        call("fpl_dice_ce_grad_ex", ptr(logits), *_src_args(truth, weight), ptr(sums), ctx.w[0], ctx.w[1], ctx.w[2],
             float(world), ptr(g), None, ptr(dlogits), n, c, spatial, ctx.w[3], n * world if world > 1 else 0, stream_ptr())
        return dlogits, None, None, None, None, None, None, None


def hard_dice_from_sums(sums, c):
    """Class-wise Dice of argmax one-hot vs ground truth (agent_seg.py:472-476) from the fused
    kernel's counters; returns a float64 CUDA tensor [C] (no sync)."""
    base = 3 * c + 2
    hi, hy, hp = sums[base:base + c], sums[base + c:base + 2 * c], sums[base + 2 * c:base + 3 * c]
    return (2.0 * hi + 1e-5) / (hy + hp + 1e-5)


class _FusedSegLoss(nn.Module):
    w_dice, w_ce = 1.0, 0.0

    def __init__(self, params=None):
        super().__init__()
        self.softmax = True if params is None else params.get('loss_softmax', True)
        # optional entropy regulariser -sum p*log2(p+1e-10)/(N*D*H*W) of agent_seg.py:353,467 ([training]
        # entropy_weight; 0 = the pinned training_all contract, where the term is commented out)
        self.w_ent = 0.0 if params is None else float(params.get('entropy_weight', 0.0) or 0.0)
        self.last = {}

    def forward(self, loss_input_dict):
        if not self.softmax and self.w_ent != 0.0:
            raise ValueError("entropy_weight needs loss_softmax = True (the term is defined on logits)")
        predict, soft_y, pix_w = _prep(loss_input_dict)
        return _DiceCE.apply(predict, soft_y, pix_w, float(self.w_dice), float(self.w_ce), float(self.w_ent),
                             0 if self.softmax else 1, self.last)

    def last_hard_dice(self):
        """Class-wise hard Dice of the last call (agent_seg.py:472-476), written by the loss-value launch."""
        h = self.last.get("hard_dice")
        if h is not None:
            return h
        s = self.last.get("sums")
        return None if s is None else hard_dice_from_sums(s, (s.numel() - 3) // 6)

    def last_entropy(self):
        """-sum p*log2(p + 1e-10) / (N*D*H*W) of the last call (needs entropy_weight != 0); CUDA float64 tensor."""
        s, v = self.last.get("sums"), self.last.get("voxels")
        return None if s is None or v is None else -s[-1] / v


class DiceLoss(_FusedSegLoss):
    w_dice, w_ce = 1.0, 0.0


class CrossEntropyLoss(_FusedSegLoss):
    w_dice, w_ce = 0.0, 1.0


class CombinedLoss(_FusedSegLoss):
    """sum_i loss_weight[i] * loss_i for ``loss_type = [..]`` (combined.py:21-39); when every term is
    a fused Dice/CE entry the sum is evaluated by ONE kernel pair."""

    def __init__(self, params, loss_dict):
        super().__init__(params)
        names = params['loss_type']
        self.loss_weight = params['loss_weight']
        assert len(names) == len(self.loss_weight)
        self.loss_list = []
        for name in names:
            if name in loss_dict:
                self.loss_list.append(loss_dict[name](params))
            else:
                raise ValueError("{0:} is not defined, or has not been added to the \
                    loss dictionary".format(name))
        self.fused = all(isinstance(l, _FusedSegLoss) and not isinstance(l, CombinedLoss) for l in self.loss_list)
        if self.fused:
            self.w_dice = sum(w * l.w_dice for w, l in zip(self.loss_weight, self.loss_list))
            self.w_ce = sum(w * l.w_ce for w, l in zip(self.loss_weight, self.loss_list))

    def forward(self, loss_input_dict):
        if self.fused:
            return super().forward(loss_input_dict)
        loss_value = 0.0
        for w, l in zip(self.loss_weight, self.loss_list):
            loss_value = loss_value + w * l(loss_input_dict)
        return loss_value
