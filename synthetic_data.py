"""Deterministic synthetic inputs for parity tests and the bench.

Everything is generated with numpy's PCG64 (platform independent), never with
torch's RNG, so the build container, the GPU box and the committed golden
vectors all see bit-identical inputs.
"""
import zlib

import numpy as np


def _rng(seed, tag=""):
    return np.random.Generator(np.random.PCG64(seed * 1000003 + zlib.crc32(tag.encode())))


def state_dict_spec(in_chns, feature_chns, class_num, num_domains):
    """Names/shapes of the 484-entry state_dict of ``UNet2D5_dsbn``
    (reference: PyMIC/pymic/net/net3d/unet2d5_dsbn.py:48-63, 131-154, 265-294)."""
    spec = []

    def conv_block(prefix, ci, co):
        spec.append((f"{prefix}.conv2d_1.weight", (co, ci, 3, 3)))
        spec.append((f"{prefix}.conv2d_1.bias", (co,)))
        spec.append((f"{prefix}.conv2d_2.weight", (co, co, 3, 3)))
        spec.append((f"{prefix}.conv2d_2.bias", (co,)))
        spec.append((f"{prefix}.conv3d_1.weight", (co, ci, 3, 3, 3)))
        spec.append((f"{prefix}.conv3d_1.bias", (co,)))
        spec.append((f"{prefix}.conv3d_2.weight", (co, co, 3, 3, 3)))
        spec.append((f"{prefix}.conv3d_2.bias", (co,)))
        for bn in ("bn2d1", "bn2d2", "bn3d1", "bn3d2"):
            for d in range(num_domains):
                spec.append((f"{prefix}.{bn}.bns.{d}.weight", (co,)))
                spec.append((f"{prefix}.{bn}.bns.{d}.bias", (co,)))
                spec.append((f"{prefix}.{bn}.bns.{d}.running_mean", (co,)))
                spec.append((f"{prefix}.{bn}.bns.{d}.running_var", (co,)))
                spec.append((f"{prefix}.{bn}.bns.{d}.num_batches_tracked", ()))
        spec.append((f"{prefix}.relu_1.weight", (1,)))
        spec.append((f"{prefix}.relu_2.weight", (1,)))

    ft = feature_chns
    chans = [in_chns] + list(ft)
    for i in range(5):
        conv_block(f"block{i}.conv", chans[i], chans[i + 1])
    for k, (c1, c2) in enumerate([(ft[4], ft[3]), (ft[3], ft[2]), (ft[2], ft[1]), (ft[1], ft[0])], start=1):
        spec.append((f"up{k}.conv2d.weight", (c2, c1, 1, 1)))
        spec.append((f"up{k}.conv2d.bias", (c2,)))
        spec.append((f"up{k}.conv3d.weight", (c2, c1, 1, 1, 1)))
        spec.append((f"up{k}.conv3d.bias", (c2,)))
        spec.append((f"up{k}.trans2d.weight", (c1, c2, 2, 2)))
        spec.append((f"up{k}.trans2d.bias", (c2,)))
        spec.append((f"up{k}.trans3d.weight", (c1, c2, 2, 2, 2)))
        spec.append((f"up{k}.trans3d.bias", (c2,)))
        conv_block(f"up{k}.conv", 2 * c2, c2)
    spec.append(("out_conv.weight", (class_num, ft[0], 1, 3, 3)))
    spec.append(("out_conv.bias", (class_num,)))
    return spec


def synth_state_dict(in_chns=1, feature_chns=(16, 32, 64, 128, 256), class_num=2,
                     num_domains=2, seed=1):
    """A full state_dict (numpy arrays) with Kaiming-uniform-like conv weights,
    BN gamma ~ U(0.5,1.5), beta ~ U(-0.2,0.2), non-trivial running stats and
    per-layer PReLU slopes, so every parameter matters in a parity check."""
    out = {}
    for name, shape in state_dict_spec(in_chns, feature_chns, class_num, num_domains):
        g = _rng(seed, name)
        leaf = name.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            out[name] = np.asarray(0, dtype=np.int64)
        elif leaf == "running_mean":
            out[name] = g.uniform(-0.1, 0.1, shape).astype(np.float32)
        elif leaf == "running_var":
            out[name] = g.uniform(0.5, 1.5, shape).astype(np.float32)
        elif ".bns." in name and leaf == "weight":
            out[name] = g.uniform(0.5, 1.5, shape).astype(np.float32)
        elif ".bns." in name and leaf == "bias":
            out[name] = g.uniform(-0.2, 0.2, shape).astype(np.float32)
        elif ".relu_" in name:
            out[name] = g.uniform(0.1, 0.4, shape).astype(np.float32)
        elif leaf == "weight":
            # transposed convs: fan_in follows torch's convention (dim 1 * receptive field)
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            out[name] = g.uniform(-bound, bound, shape).astype(np.float32)
        else:  # conv bias
            out[name] = g.uniform(-0.05, 0.05, shape).astype(np.float32)
    return out


def synth_image(n, in_chns, shape, seed=1, tag="img"):
    """fp32 ~ N(0,1) volumes (what NormalizeWithMeanStd yields; reference
    PyMIC/pymic/transform/normalize.py:58-60)."""
    g = _rng(seed, tag)
    return g.standard_normal((n, in_chns) + tuple(shape)).astype(np.float32)


def synth_label(n, class_num, shape, seed=1, tag="lab"):
    """uint8 label maps: background + 1-3 random ellipsoids per foreground class."""
    g = _rng(seed, tag)
    D, H, W = shape
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    lab = np.zeros((n, D, H, W), dtype=np.uint8)
    for i in range(n):
        for c in range(1, class_num):
            for _ in range(int(g.integers(1, 4))):
                ctr = [g.uniform(0.2, 0.8) * s for s in (D, H, W)]
                rad = [max(1.5, g.uniform(0.08, 0.22) * s) for s in (D, H, W)]
                m = (((zz - ctr[0]) / rad[0]) ** 2 + ((yy - ctr[1]) / rad[1]) ** 2
                     + ((xx - ctr[2]) / rad[2]) ** 2) <= 1.0
                lab[i][m] = c
    return lab


def one_hot(lab, class_num):
    """[N,D,H,W] uint8 -> [N,C,D,H,W] fp32 (reference LabelToProbability,
    PyMIC/pymic/transform/label_convert.py:82-88)."""
    return np.stack([(lab == c) for c in range(class_num)], axis=1).astype(np.float32)


def synth_pixel_weight(lab, seed=1, tag="pw"):
    """1.0 with a one-voxel shell around the foreground set to 0.5 (the
    agreement-map look of data/get_pixel_weight.py), then folded with a
    per-image weight ~ U(0.01,1.01) exactly as NiftyDataset.set_weight_ does
    (PyMIC/pymic/io/nifty_dataset.py:165-168).  Returns ([N,1,D,H,W] fp32, [N] f64)."""
    g = _rng(seed, tag)
    n = lab.shape[0]
    fg = lab > 0
    dil = fg.copy()
    for ax in (1, 2, 3):
        dil |= np.roll(fg, 1, axis=ax) | np.roll(fg, -1, axis=ax)
    w = np.ones(lab.shape, dtype=np.float32)
    w[dil & ~fg] = 0.5
    img_w = g.uniform(0.01, 1.01, n)
    w[w < 1] = 0
    w = w * img_w.astype(np.float32)[:, None, None, None]
    return w[:, None].astype(np.float32), img_w
