"""The filtered-pseudo-label arithmetic on the device (csrc/filter.cu), plus the tiny host-side
ordering/weight-map steps.

References: pseudo labels agent_seg.py:1045-1050; MC-dropout image uncertainty agent_seg.py:897-931;
ordering agent_seg.py:954-960; agreement weight data/get_pixel_weight.py:21-26; image-weight folding
io/nifty_dataset.py:165-168; image-weight map recovered from dataset/weight/cyc121_vst1s-gan.npy +
config_dual/data_vs/train_vs_t1s_wi+wp.csv.
"""
import ctypes

import numpy as np
import torch

from .ops import call, ptr, stream_ptr


def pseudo_label(logits):
    """fp32 logits [B,C,D,H,W] (CUDA) -> uint8 labels [B,D,H,W]."""
    logits = logits.float().contiguous()
    b, c = logits.shape[:2]
    spatial = logits.numel() // (b * c)
    out = torch.empty((b,) + tuple(logits.shape[2:]), dtype=torch.uint8, device=logits.device)
    call("fpl_argmax_label", ptr(logits), ptr(out), b, c, spatial, stream_ptr())
    return out


def max_mc_passes():
    """Largest K of ``mc_uncertainty`` (the reference hard-codes 6, agent_seg.py:898); callers validate
    ``fpl_mc_passes`` against it before running the K forwards."""
    from . import lib as _lib
    return int(_lib.load().fpl_mc_uncertainty_max_passes())


def mc_uncertainty(logit_passes, want_map=False):
    """K MC-dropout logits maps of one volume (each [1,C,D,H,W] CUDA fp32).
    Returns (stats, umap): stats is a float64 CUDA tensor [2] = (sum of variances, boundary count)
    -- no host sync here; ``finish_uncertainty`` turns it into the reference's ``uncer_one``."""
    passes = [p.float().contiguous() for p in logit_passes]
    k = len(passes)
    if not 1 <= k <= max_mc_passes():
        raise ValueError("mc_uncertainty: %d passes not in [1, %d]" % (k, max_mc_passes()))
    _b, c = passes[0].shape[:2]
    if passes[0].shape[0] != 1:
        raise ValueError("mc_uncertainty expects one volume per call (batch 1), as the reference test loader")
    spatial = passes[0].numel() // c
    out = torch.zeros(2, dtype=torch.float64, device=passes[0].device)
    umap = torch.empty(passes[0].shape[2:], dtype=torch.float32, device=passes[0].device) if want_map else None
    arr = (ctypes.c_void_p * k)(*[p.data_ptr() for p in passes])
    call("fpl_mc_uncertainty", arr, k, c, spatial, ptr(out), ptr(umap), stream_ptr())
    return out, umap


def finish_uncertainty(stats):
    """(vars, boundary) -> uncer_one: 1 if boundary < 50 else vars/boundary (agent_seg.py:926-929).
    ``stats`` may be a CUDA tensor (one D2H read of 16 bytes) or a (vars, boundary) pair."""
    v, b = [float(t) for t in (stats.tolist() if torch.is_tensor(stats) else stats)]
    boundary = int(round(b))
    return 1 if boundary < 50 else np.float64(np.float32(v)) / boundary


def agreement_weight(logits_tgt, logits_src, image_weight=None):
    """Target-domain and fake-source logits of one volume -> (label_tgt u8, label_src u8,
    weight fp32, n_disagree int64 tensor).  weight = 1 - 0.5*[a != b]; with ``image_weight`` it is
    folded as NiftyDataset.set_weight_ does ((w<1 ? 0 : w) * image_weight)."""
    lt, ls = logits_tgt.float().contiguous(), logits_src.float().contiguous()
    if lt.shape != ls.shape:
        raise ValueError("the two logits maps differ in shape")
    b, c = lt.shape[:2]
    if b != 1:
        raise ValueError("agreement_weight expects one volume per call")
    spatial = lt.numel() // c
    shp = tuple(lt.shape[2:])
    la = torch.empty(shp, dtype=torch.uint8, device=lt.device)
    lb = torch.empty(shp, dtype=torch.uint8, device=lt.device)
    w = torch.empty(shp, dtype=torch.float32, device=lt.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=lt.device)
    fold = 0 if image_weight is None else 1
    call("fpl_agree_weight", ptr(lt), ptr(ls), ptr(la), ptr(lb), ptr(w), fold, float(image_weight or 0.0), ptr(cnt),
         c, spatial, stream_ptr())
    return la, lb, w, cnt


def sort_uncertainty(uncertainty_by_name):
    """{name: [uncer_one]} -> ascending [([value], name)], ties broken by the name string (python
    tuple ordering, exactly agent_seg.py:957-958).  ~100 items: stays on the host."""
    return sorted(zip(uncertainty_by_name.values(), uncertainty_by_name.keys()), reverse=False)


def image_weights(sorted_values):
    """w = 1.01 - (u - u_min)/(u_max* - u_min); sentinel u == 1 -> 0.01 (u_max* = largest non-sentinel)."""
    u = np.asarray([float(v) for v in sorted_values], np.float64)
    real = u[u != 1.0]
    lo, hi = real.min(), real.max()
    w = 1.01 - (u - lo) / (hi - lo)
    w[u == 1.0] = 0.01
    return w
