"""Oracle: functional fp32 restatement of ``UNet2D5_dsbn`` (test infrastructure).

Follows PyMIC/pymic/net/net3d/unet2d5_dsbn.py (ConvBlockND :65-83, DownBlock
:108-129, UpBlock :156-188, UNet2D5_dsbn.forward :296-309) and
PyMIC/pymic/net_run_dsbn/dsbn.py:54-57 (``bns[domain_label[0]]``).

The network is evaluated straight from a ``state_dict`` (dict name -> tensor)
with ``torch.nn.functional`` ops on CPU in fp32, so gradients come from
autograd when the tensors require grad.  BatchNorm running statistics are
updated in place in ``state`` for the selected domain exactly as
``nn.BatchNorm3d(momentum=0.1)`` does.
"""
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# bf16 data-path emulation (``bf16=True``): the SAME fp32 restatement with the rounding points of
# the CUDA path made explicit, so that kernel parity (tight tolerance against this) is separated
# from the precision effect of bf16 storage (this against the plain fp32 oracle).  Rounding
# points of fplplus_b200: conv / transposed-conv outputs and activations are STORED in bf16;
# tensor-core convs see bf16-rounded weights; BatchNorm statistics come from the fp32 conv
# results BEFORE rounding; gradients wrt conv inputs (dgrad outputs) and wrt conv outputs (the
# BatchNorm backward output) are stored in bf16; every accumulation is fp32.
# ---------------------------------------------------------------------------------------------
def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _RoundFwd(torch.autograd.Function):       # stored in bf16; gradient passes unchanged
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):       # identity forward; the gradient is stored in bf16
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


def _wq(w):
    """bf16-rounded weights for the tensor-core convs, straight-through for dW."""
    return w + (_r(w) - w).detach()


class _BNStored(torch.autograd.Function):
    """BatchNorm over a conv result that is stored in bf16: statistics from the fp32 result,
    normalisation (and the backward's xhat) from the rounded copy, dy stored in bf16
    (csrc/dsbn.cu: dsbn_finalize_kernel, dsbn_act_fwd_kernel, dsbn_act_bwd_kernel)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, rm, rv, training, momentum, eps):
        dims = [0] + list(range(2, y.dim()))
        shape = [1, -1] + [1] * (y.dim() - 2)
        if training:
            n = y.numel() // y.shape[1]
            yd = y.double()
            mean = yd.mean(dims)
            var = ((yd * yd).mean(dims) - mean * mean).clamp_(min=0)
            invstd = (1.0 / torch.sqrt(var + eps)).float()
            with torch.no_grad():
                rm.mul_(1 - momentum).add_(momentum * mean.float())
                rv.mul_(1 - momentum).add_(momentum * (var * (n / max(n - 1, 1))).float())
            mean = mean.float()
        else:
            mean, invstd = rm.clone(), 1.0 / torch.sqrt(rv + eps)
        scale = gamma * invstd
        shift = beta - mean * scale
        yq = _r(y)
        ctx.save_for_backward(yq, mean, invstd, scale)
        ctx.training, ctx.dims, ctx.shape = training, dims, shape
        return yq * scale.view(shape) + shift.view(shape)

    @staticmethod
    def backward(ctx, dz):
        yq, mean, invstd, scale = ctx.saved_tensors
        dims, shape = ctx.dims, ctx.shape
        xhat = (yq - mean.view(shape)) * invstd.view(shape)
        dbeta = dz.sum(dims)
        dgamma = (dz * xhat).sum(dims)
        if ctx.training:
            n = dz.numel() // dz.shape[1]
            dy = scale.view(shape) * (dz - (dbeta / n).view(shape) - xhat * (dgamma / n).view(shape))
        else:
            dy = scale.view(shape) * dz
        return _r(dy), dgamma, dbeta, None, None, None, None, None


def _bn(state, prefix, x, domain, training, bf16=False):
    if bf16:
        p = f"{prefix}.bns.{domain}"
        y = _BNStored.apply(x, state[p + ".weight"], state[p + ".bias"], state[p + ".running_mean"],
                            state[p + ".running_var"], training, 0.1, 1e-5)
        if training:
            state[p + ".num_batches_tracked"] += 1
        return y
    return _bn_fp32(state, prefix, x, domain, training)


def _bn_fp32(state, prefix, x, domain, training):
    p = f"{prefix}.bns.{domain}"
    rm, rv = state[p + ".running_mean"], state[p + ".running_var"]
    y = F.batch_norm(x, rm, rv, state[p + ".weight"], state[p + ".bias"],
                     training=training, momentum=0.1, eps=1e-5)
    if training:
        state[p + ".num_batches_tracked"] += 1
    return y


def _dropout(x, p, key, training, masks):
    """Element-wise dropout.  ``masks`` (dict key -> bool keep-mask shaped like x)
    makes the draw explicit so a CUDA kernel can consume the same mask."""
    if not training or p <= 0.0:
        return x
    if masks is not None and key in masks:
        return x * masks[key].to(x.dtype) / (1.0 - p)
    return F.dropout(x, p, True)


def conv_block(state, prefix, x, domain, dim, p_drop, bn_training, drop_training, masks=None, bf16=False,
               stem=False):
    conv = F.conv2d if dim == 2 else F.conv3d
    cn, bn = ("conv2d", "bn2d") if dim == 2 else ("conv3d", "bn3d")
    for k in (1, 2):
        w, b = state[f"{prefix}.{cn}_{k}.weight"], state[f"{prefix}.{cn}_{k}.bias"]
        if bf16:
            # the stem is evaluated from hi/lo bf16 pairs of the image and of its weights: fp32-accurate
            first = stem and k == 1
            x = conv(x if first else _RoundBwd.apply(x), w if first else _wq(w), b, padding=1)
        else:
            x = conv(x, w, b, padding=1)
        x = _bn(state, f"{prefix}.{bn}{k}", x, domain, bn_training, bf16)
        x = F.prelu(x, state[f"{prefix}.relu_{k}.weight"])
        if k == 1:
            x = _dropout(x, p_drop, prefix, drop_training, masks)
        if bf16:
            x = _RoundFwd.apply(x)
    return x


def _to2d(x):
    n, c, d, h, w = x.shape
    return x.transpose(1, 2).reshape(n * d, c, h, w), (n, d)


def _to3d(x, nd):
    n, d = nd
    return x.reshape((n, d) + tuple(x.shape[1:])).transpose(1, 2)


def forward(state, x, domain, params, bn_training=False, drop_training=None, masks=None, bf16=False):
    """logits = UNet2D5_dsbn(params)(x, domain_label=domain*ones(N)).

    ``params`` is the reference's ``config['network']`` dict.  ``drop_training``
    defaults to ``bn_training`` (module.train()); FPL test-time dropout is
    ``bn_training=False, drop_training=True`` (agent_seg.py:843-852).  ``bf16=True`` emulates
    the storage roundings of the CUDA path (see the top of this file)."""
    if drop_training is None:
        drop_training = bn_training
    dims, drop, bilinear = params["conv_dims"], params["dropout"], params["bilinear"]
    skips = []
    h = x
    for i in range(5):
        pre = f"block{i}.conv"
        if dims[i] == 2:
            h2, nd = _to2d(h)
            o = conv_block(state, pre, h2, domain, 2, drop[i], bn_training, drop_training, masks, bf16, i == 0)
            od = F.max_pool2d(o, 2, 2) if i < 4 else None
            o = _to3d(o, nd)
            od = _to3d(od, nd) if od is not None else None
        else:
            o = conv_block(state, pre, h, domain, 3, drop[i], bn_training, drop_training, masks, bf16, i == 0)
            od = F.max_pool3d(o, 2, 2) if i < 4 else None
        skips.append(o)
        h = od
    h = skips[4]
    for k, lvl in zip((1, 2, 3, 4), (3, 2, 1, 0)):
        pre = f"up{k}"
        skip = skips[lvl]
        if dims[lvl] == 2:
            h2, nd = _to2d(h)
            s2, _ = _to2d(skip)
            if bilinear:
                h2 = F.conv2d(h2, state[pre + ".conv2d.weight"], state[pre + ".conv2d.bias"])
                h2 = F.interpolate(h2, scale_factor=2, mode="bilinear", align_corners=True)
            else:
                h2 = F.conv_transpose2d(_RoundBwd.apply(h2) if bf16 else h2,
                                        _wq(state[pre + ".trans2d.weight"]) if bf16 else state[pre + ".trans2d.weight"],
                                        state[pre + ".trans2d.bias"], stride=2)
                h2 = _RoundFwd.apply(h2) if bf16 else h2
            cat = torch.cat([s2, h2], dim=1)
            o = conv_block(state, pre + ".conv", cat, domain, 2, drop[lvl], bn_training, drop_training, masks, bf16)
            h = _to3d(o, nd)
        else:
            if bilinear:
                h = F.conv3d(h, state[pre + ".conv3d.weight"], state[pre + ".conv3d.bias"])
                h = F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True)
            else:
                h = F.conv_transpose3d(_RoundBwd.apply(h) if bf16 else h,
                                       _wq(state[pre + ".trans3d.weight"]) if bf16 else state[pre + ".trans3d.weight"],
                                       state[pre + ".trans3d.bias"], stride=2)
                h = _RoundFwd.apply(h) if bf16 else h
            cat = torch.cat([skip, h], dim=1)
            h = conv_block(state, pre + ".conv", cat, domain, 3, drop[lvl], bn_training, drop_training, masks, bf16)
    if bf16:     # the head runs on the tensor cores too: bf16 weights, dlogits stored in bf16 for dgrad / wgrad
        return _RoundBwd.apply(F.conv3d(_RoundBwd.apply(h), _wq(state["out_conv.weight"]), state["out_conv.bias"],
                                        padding=(0, 1, 1)))
    return F.conv3d(h, state["out_conv.weight"], state["out_conv.bias"], padding=(0, 1, 1))


def to_torch_state(np_state, requires_grad=False):
    out = {}
    for k, v in np_state.items():
        t = torch.from_numpy(v.copy()) if hasattr(v, "dtype") and not torch.is_tensor(v) else v.clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out
