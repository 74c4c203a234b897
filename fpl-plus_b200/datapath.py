"""Device data path of the training inputs.

PyMIC feeds the step from 16 CPU workers that decode float64 NIfTI, normalise, pad, crop, flip and build an fp32 one-hot
per sample (io/nifty_dataset.py:171-218, transform/{normalize,pad,crop,flip,label_convert}.py); that pipeline cannot
feed a B200 at ~10^9 voxels/s.  Here the (already normalised and padded) volumes of a domain live in HBM for the whole
run and one gather kernel per step (csrc/datapath.cu) cuts and flips the batch:

* ``RandomCrop`` with foreground focus (transform/crop.py:170-244) and ``RandomFlip`` (flip.py:14-62): the crop origin
  and the flip axes are drawn on the host with ``random.Random`` in the SAME order of draws as the reference transforms
  (per sample: one randint per axis with a non-zero margin, the foreground coin, the three bounding-box draws; then the
  width / height / depth coins), so a run seeded like the reference's workers makes the same decisions;
* ``LabelToProbability`` (label_convert.py:82-88) and ``NiftyDataset.set_weight_`` (nifty_dataset.py:165-168) are NOT
  applied here: the batch carries the uint8 label map, the uint8 agreement code (0/1/2 = weight 0/0.5/1) and the
  per-sample image weight, and the loss kernels (fpl_dice_ce_*_ex) expand them per voxel.

The batch dicts it yields have the keys ``SegmentationAgent.train_step`` consumes: ``image`` fp32 [N,C,D,H,W], ``label``
uint8 [N,D,H,W], ``pixel_weight`` uint8 [N,1,D,H,W], ``image_weight`` fp32 [N] (device), ``names``.
"""
import random
import struct

import numpy as np
import torch

from .ops import call, ptr, stream_ptr


def normalize_with_mean_std(image, mask_nonzero=False):
    """NormalizeWithMeanStd (transform/normalize.py:36-62) per channel: (x - mean) / std; with ``mask_nonzero`` the
    statistics come from the voxels > 0 (the rest is left untouched apart from the affine map, as in the reference's
    default ``random_fill = False`` path)."""
    img = np.asarray(image, np.float32).copy()
    for c in range(img.shape[0]):
        ch = img[c]
        sel = ch[ch > 0] if mask_nonzero and (ch > 0).any() else ch
        img[c] = (ch - sel.mean()) / max(float(sel.std()), 1e-8)
    return img


def pad_to(volume, out_size, mode="reflect"):
    """Pad (transform/pad.py:60-90): centre-pad the trailing three axes up to ``out_size``."""
    shape = volume.shape[-3:]
    margin = [max(0, out_size[i] - shape[i]) for i in range(3)]
    if not any(margin):
        return volume
    lo = [m // 2 for m in margin]
    hi = [m - l for m, l in zip(margin, lo)]
    pad = [(0, 0)] * (volume.ndim - 3) + list(zip(lo, hi))
    return np.pad(volume, pad, mode if all(s > 1 for s in shape) else "edge")


def code_from_pixel_weight(w):
    """Agreement weight map {0, 0.5, 1} (data/get_pixel_weight.py:21-26) -> uint8 code {0, 1, 2}."""
    return np.rint(np.asarray(w, np.float32) * 2.0).astype(np.uint8)


class DevicePatchSampler(object):
    """Resident volumes of ONE domain + per-step RandomCrop / RandomFlip on the device.

    ``volumes``: list of dicts with ``image`` [C,D,H,W] fp32 (normalised), optional ``label`` [D,H,W] integer,
    optional ``pixel_weight`` [D,H,W] in {0, 0.5, 1} (or ``code`` uint8), optional ``image_weight`` float, ``name``.
    Iterating yields batch dicts forever (one epoch = a random permutation of the volumes, like a shuffling DataLoader
    with ``drop_last``)."""

    def __init__(self, volumes, patch, batch_size, device, fg_focus=True, fg_ratio=0.5, mask_label=(1,),
                 flip=(False, True, True), seed=1, rank=0, world=1):
        self.patch = tuple(int(p) for p in patch)
        self.batch = int(batch_size)
        self.device = torch.device(device)
        self.fg_focus, self.fg_ratio, self.mask_label = fg_focus, fg_ratio, tuple(mask_label)
        self.flip_depth, self.flip_height, self.flip_width = flip
        self.rng = random.Random(int(seed) * 1000003 + rank)
        self.rank, self.world = rank, world
        self.vols = []
        for v in volumes:
            img = pad_to(np.asarray(v["image"], np.float32), self.patch)
            if img.ndim == 3:
                img = img[None]
            ent = {"name": v.get("name", "vol%d" % len(self.vols)), "shape": img.shape[1:],
                   "image": torch.from_numpy(np.ascontiguousarray(img)).to(self.device),
                   "label": None, "code": None, "bbox": None, "image_weight": float(v.get("image_weight", 1.0))}
            if v.get("label") is not None:
                lab = pad_to(np.asarray(v["label"]).astype(np.uint8), self.patch, "constant")
                ent["label"] = torch.from_numpy(np.ascontiguousarray(lab)).to(self.device)
                mask = np.isin(lab, self.mask_label)
                if mask.any():
                    idx = np.nonzero(mask)
                    # get_ND_bounding_box (util/image_process.py:8-35), margin 0: bb_max is max + 1
                    ent["bbox"] = ([int(i.min()) for i in idx], [int(i.max()) + 1 for i in idx])
            code = v.get("code")
            if code is None and v.get("pixel_weight") is not None:
                code = code_from_pixel_weight(v["pixel_weight"])
            if code is not None:
                code = pad_to(np.asarray(code, np.uint8), self.patch, "constant")
                ent["code"] = torch.from_numpy(np.ascontiguousarray(code)).to(self.device)
            self.vols.append(ent)
        if not self.vols:
            raise ValueError("DevicePatchSampler needs at least one volume")
        self.in_chns = int(self.vols[0]["image"].shape[0])
        self.has_label = all(v["label"] is not None for v in self.vols)
        self.has_code = any(v["code"] is not None for v in self.vols)
        self._order, self._pos = [], 0
        # double-buffered outputs + pinned parameter tables: the batch of step i+1 is gathered while step i trains
        n, (pd, ph, pw) = self.batch, self.patch
        self._out = [{"image": torch.empty((n, self.in_chns, pd, ph, pw), dtype=torch.float32, device=self.device),
                      "label": torch.empty((n, pd, ph, pw), dtype=torch.uint8, device=self.device) if self.has_label else None,
                      "code": torch.empty((n, 1, pd, ph, pw), dtype=torch.uint8, device=self.device) if self.has_code else None,
                      "iw": torch.empty(n, dtype=torch.float32, device=self.device),
                      "rows": torch.empty(64 * n, dtype=torch.uint8, device=self.device),
                      "rows_host": torch.empty(64 * n, dtype=torch.uint8).pin_memory() if self.device.type == "cuda" else None,
                      "iw_host": torch.empty(n, dtype=torch.float32).pin_memory() if self.device.type == "cuda" else None,
                      "copied": None}
                     for _ in range(2)]
        self._slot = 0
        self.last_params = None

    # -- the reference's random decisions ------------------------------------------------------------------
    def _next_volume(self):
        if self._pos >= len(self._order):
            idx = list(range(self.rank, len(self.vols), self.world)) or list(range(len(self.vols)))
            self.rng.shuffle(idx)
            self._order, self._pos = idx, 0
        i = self._order[self._pos]
        self._pos += 1
        return i

    def draw(self, vol):
        """(crop origin, flip bits) of one sample: transform/crop.py:213-234 then flip.py:38-47."""
        shape, out = vol["shape"], self.patch
        margin = [shape[i] - out[i] for i in range(3)]
        crop_min = [0 if m == 0 else self.rng.randint(0, m) for m in margin]
        if self.fg_focus and self.rng.random() < self.fg_ratio:
            if vol["bbox"] is None:
                bb_min, bb_max = [0, 0, 0], list(shape)
            else:
                bb_min, bb_max = vol["bbox"]
            crop_min = [self.rng.randint(bb_min[i], bb_max[i]) - int(out[i] / 2) for i in range(3)]
            crop_min = [max(0, c) for c in crop_min]
            crop_min = [min(crop_min[i], shape[i] - out[i]) for i in range(3)]
        flip = 0
        if self.flip_width and self.rng.random() > 0.5:
            flip |= 4
        if self.flip_height and self.rng.random() > 0.5:
            flip |= 2
        if self.flip_depth and self.rng.random() > 0.5:
            flip |= 1
        return crop_min, flip

    # -- one batch ---------------------------------------------------------------------------------------------
    def next_batch(self):
        out = self._out[self._slot]
        self._slot ^= 1
        rows, names, params = bytearray(), [], []
        iw = []
        for _ in range(self.batch):
            v = self.vols[self._next_volume()]
            (d0, h0, w0), flip = self.draw(v)
            D, H, W = v["shape"]
            rows += struct.pack("<3Q7i3i", v["image"].data_ptr(), v["label"].data_ptr() if v["label"] is not None else 0,
                                v["code"].data_ptr() if v["code"] is not None else 0, D, H, W, d0, h0, w0, flip, 0, 0, 0)
            names.append(v["name"])
            iw.append(v["image_weight"])
            params.append((v["name"], (d0, h0, w0), flip))
        self.last_params = params
        if out["rows_host"] is not None:
            if out["copied"] is not None:
                out["copied"].synchronize()          # the pinned tables of this slot were last read two batches ago
            out["rows_host"].copy_(torch.frombuffer(rows, dtype=torch.uint8))
            out["rows"].copy_(out["rows_host"], non_blocking=True)
            out["iw_host"].copy_(torch.tensor(iw, dtype=torch.float32))
            out["iw"].copy_(out["iw_host"], non_blocking=True)
            out["copied"] = torch.cuda.Event()
            out["copied"].record()
        else:
            out["rows"].copy_(torch.frombuffer(rows, dtype=torch.uint8))
            out["iw"].copy_(torch.tensor(iw, dtype=torch.float32))
        pd, ph, pw = self.patch
        call("fpl_gather_patches", ptr(out["rows"]), self.batch, self.in_chns, pd, ph, pw, ptr(out["image"]), ptr(out["label"]),
             ptr(out["code"]), stream_ptr())
        batch = {"image": out["image"], "names": names}
        if self.has_label:
            batch["label"] = out["label"]
        if self.has_code:
            batch["pixel_weight"] = out["code"]
            batch["image_weight"] = out["iw"]
        return batch

    def __iter__(self):
        return self

    def __next__(self):
        return self.next_batch()

    def __len__(self):
        return max(1, len(self.vols) // (self.batch * self.world))
