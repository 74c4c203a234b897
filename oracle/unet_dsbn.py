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


def _bn(state, prefix, x, domain, training):
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


def conv_block(state, prefix, x, domain, dim, p_drop, bn_training, drop_training, masks=None):
    if dim == 2:
        x = F.conv2d(x, state[prefix + ".conv2d_1.weight"], state[prefix + ".conv2d_1.bias"], padding=1)
        x = _bn(state, prefix + ".bn2d1", x, domain, bn_training)
        x = F.prelu(x, state[prefix + ".relu_1.weight"])
        x = _dropout(x, p_drop, prefix, drop_training, masks)
        x = F.conv2d(x, state[prefix + ".conv2d_2.weight"], state[prefix + ".conv2d_2.bias"], padding=1)
        x = _bn(state, prefix + ".bn2d2", x, domain, bn_training)
        x = F.prelu(x, state[prefix + ".relu_2.weight"])
    else:
        x = F.conv3d(x, state[prefix + ".conv3d_1.weight"], state[prefix + ".conv3d_1.bias"], padding=1)
        x = _bn(state, prefix + ".bn3d1", x, domain, bn_training)
        x = F.prelu(x, state[prefix + ".relu_1.weight"])
        x = _dropout(x, p_drop, prefix, drop_training, masks)
        x = F.conv3d(x, state[prefix + ".conv3d_2.weight"], state[prefix + ".conv3d_2.bias"], padding=1)
        x = _bn(state, prefix + ".bn3d2", x, domain, bn_training)
        x = F.prelu(x, state[prefix + ".relu_2.weight"])
    return x


def _to2d(x):
    n, c, d, h, w = x.shape
    return x.transpose(1, 2).reshape(n * d, c, h, w), (n, d)


def _to3d(x, nd):
    n, d = nd
    return x.reshape((n, d) + tuple(x.shape[1:])).transpose(1, 2)


def forward(state, x, domain, params, bn_training=False, drop_training=None, masks=None):
    """logits = UNet2D5_dsbn(params)(x, domain_label=domain*ones(N)).

    ``params`` is the reference's ``config['network']`` dict.  ``drop_training``
    defaults to ``bn_training`` (module.train()); FPL test-time dropout is
    ``bn_training=False, drop_training=True`` (agent_seg.py:843-852)."""
    if drop_training is None:
        drop_training = bn_training
    dims, drop, bilinear = params["conv_dims"], params["dropout"], params["bilinear"]
    skips = []
    h = x
    for i in range(5):
        pre = f"block{i}.conv"
        if dims[i] == 2:
            h2, nd = _to2d(h)
            o = conv_block(state, pre, h2, domain, 2, drop[i], bn_training, drop_training, masks)
            od = F.max_pool2d(o, 2, 2) if i < 4 else None
            o = _to3d(o, nd)
            od = _to3d(od, nd) if od is not None else None
        else:
            o = conv_block(state, pre, h, domain, 3, drop[i], bn_training, drop_training, masks)
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
                h2 = F.conv_transpose2d(h2, state[pre + ".trans2d.weight"], state[pre + ".trans2d.bias"], stride=2)
            cat = torch.cat([s2, h2], dim=1)
            o = conv_block(state, pre + ".conv", cat, domain, 2, drop[lvl], bn_training, drop_training, masks)
            h = _to3d(o, nd)
        else:
            if bilinear:
                h = F.conv3d(h, state[pre + ".conv3d.weight"], state[pre + ".conv3d.bias"])
                h = F.interpolate(h, scale_factor=2, mode="trilinear", align_corners=True)
            else:
                h = F.conv_transpose3d(h, state[pre + ".trans3d.weight"], state[pre + ".trans3d.bias"], stride=2)
            cat = torch.cat([skip, h], dim=1)
            h = conv_block(state, pre + ".conv", cat, domain, 3, drop[lvl], bn_training, drop_training, masks)
    return F.conv3d(h, state["out_conv.weight"], state["out_conv.bias"], padding=(0, 1, 1))


def to_torch_state(np_state, requires_grad=False):
    out = {}
    for k, v in np_state.items():
        t = torch.from_numpy(v.copy()) if hasattr(v, "dtype") and not torch.is_tensor(v) else v.clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out
