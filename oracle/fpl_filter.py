"""Oracle: the filtered-pseudo-label arithmetic in numpy (test infrastructure).

References (all /root/reference):
  * plain pseudo labels: net_run_dsbn/agent_seg.py:1045-1050 (scipy softmax -> argmax -> uint8)
  * MC-dropout image uncertainty: agent_seg.py:897-931
  * ascending (value, name) ordering and .npy layout: agent_seg.py:954-960
  * dual-domain agreement pixel weight: data/get_pixel_weight.py:21-26
  * folding image weight into pixel weight: PyMIC/pymic/io/nifty_dataset.py:165-168
  * image-weight map: the script is missing from the tree (README.md:74); the
    affine map below is recovered from the two shipped artefacts and reproduces
    all 100 rows (tests/golden/fpl_image_weights.json, max abs err ~1e-16).
"""
import numpy as np
from scipy.special import softmax as _softmax


def pseudo_label(logits):
    """[B,C,D,H,W] fp32 logits -> uint8 [B,D,H,W] (agent_seg.py:1049-1050)."""
    prob = _softmax(logits, axis=1)
    return np.asarray(np.argmax(prob, axis=1), np.uint8)


def mc_uncertainty(logit_passes):
    """K MC-dropout passes of one volume, each [1,C,D,H,W] fp32 logits.
    Returns dict(vars, boundary, uncer_one, uncertainty_map, hards) following
    agent_seg.py:911-929 line by line (fp32 maps, population variance over the
    K passes summed over classes and voxels, class-1 mean, thresholds 0.01/50)."""
    maps, hards = None, None
    for i, pred in enumerate(logit_passes):
        prob = _softmax(np.asarray(pred, np.float32), axis=1)
        hard = np.asarray(np.argmax(prob, axis=1), np.uint8)
        if i == 0:
            maps, hards = prob, hard
        else:
            maps = np.concatenate((maps, prob), axis=0)
            hards = np.concatenate((hards, hard), axis=0)
    vars_ = maps.var(axis=0).sum()
    means = np.mean(maps[:, 1], axis=0)
    uncertainty = -1.0 * (means * np.log(means + 1e-6))
    boundary = np.where(uncertainty > 0.01, 1, 0).sum()
    uncer_one = 1 if boundary < 50 else vars_ / boundary
    return {"vars": vars_, "boundary": int(boundary), "uncer_one": uncer_one,
            "uncertainty_map": uncertainty, "hards": hards, "means": means}


def sort_uncertainty(uncertainty_by_name):
    """{name: [uncer_one]} -> ascending list of ([value], name) tuples, ties broken by
    name (python tuple ordering) -- agent_seg.py:957-958."""
    pairs = list(zip(uncertainty_by_name.values(), uncertainty_by_name.keys()))
    return sorted(pairs, reverse=False)


def agreement_weight(label_target, label_fake_source):
    """uint8 labels of the target-domain pass and the fake-source pass ->
    float64 weight, 1 where they agree and 0.5 where they differ
    (data/get_pixel_weight.py:21-26; uint8 arithmetic kept as written, so it is
    only meaningful for binary labels)."""
    a = np.asarray(label_target)
    b = np.asarray(label_fake_source)
    both = a + b
    both[both > 1] = 1
    and_arr = b * a
    sub = both - and_arr
    return np.ones_like(sub) - sub * 0.5


def agreement_weight_multiclass(label_target, label_fake_source):
    """w = 1 - 0.5*[a != b]: identical to agreement_weight on {0,1} labels, and
    the generalisation used for class_num > 2."""
    return 1.0 - 0.5 * (np.asarray(label_target) != np.asarray(label_fake_source))


def set_weight_(img_weight, pixel_weight):
    """NiftyDataset.set_weight_ (nifty_dataset.py:165-168): voxels with weight < 1
    are zeroed, the rest multiplied by the image weight."""
    pw = np.array(pixel_weight, copy=True)
    pw[pw < 1] = 0
    return pw * img_weight


def image_weights(sorted_uncertainty):
    """Ascending uncertainties (floats; sentinel 1 = 'fewer than 50 boundary voxels')
    -> image weights.  w = 1.01 - (u - u_min)/(u_max* - u_min), u_max* the largest
    non-sentinel value; sentinel rows get 0.01."""
    u = np.asarray([float(v) for v in sorted_uncertainty], np.float64)
    real = u[u != 1.0]
    u_min, u_max = real.min(), real.max()
    w = 1.01 - (u - u_min) / (u_max - u_min)
    w[u == 1.0] = 0.01
    return w
