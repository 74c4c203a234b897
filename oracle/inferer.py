"""Oracle: sliding-window + flip-TTA inference (test infrastructure).

Follows PyMIC/pymic/net_run_dsbn/infer_func.py: window enumeration :75-85
(w outer, h, d inner; start = min(k*stride, dim-win)), clamping :65-69, whole
image shortcut :71-73, accumulate/count/divide :96-112, 4-flip TTA averaging
logits :199-219.  The reference's probe forward on a ones tensor (:92-93) only
counts outputs; it changes no result for a deterministic model and is not
restated (it does consume dropout RNG in MC mode, which no parity test can
observe without the same generator).
"""
import torch


def window_starts(img_shape, window_size, window_stride):
    win = list(window_size)
    stride = list(window_stride)
    for d in range(3):
        if win[d] is None or win[d] > img_shape[d]:
            win[d] = img_shape[d]
        if stride[d] is None or stride[d] > win[d]:
            stride[d] = win[d]
    if all(win[d] >= img_shape[d] for d in range(3)):
        return None, win
    starts = []
    for w in range(0, img_shape[2], stride[2]):
        w0 = min(w, img_shape[2] - win[2])
        for h in range(0, img_shape[1], stride[1]):
            h0 = min(h, img_shape[1] - win[1])
            for d in range(0, img_shape[0], stride[0]):
                d0 = min(d, img_shape[0] - win[0])
                starts.append((d0, h0, w0))
    return starts, win


def sliding_window(model, image, class_num, window_size, window_stride):
    """``model(patch) -> logits``; image [B,Cin,D,H,W]."""
    starts, win = window_starts(list(image.shape[2:]), window_size, window_stride)
    if starts is None:
        return model(image)
    b = image.shape[0]
    out = torch.zeros([b, class_num] + list(image.shape[2:]), dtype=image.dtype)
    cnt = torch.zeros_like(out)
    for d0, h0, w0 in starts:
        sl = (slice(None), slice(None), slice(d0, d0 + win[0]), slice(h0, h0 + win[1]), slice(w0, w0 + win[2]))
        out[sl] += model(image[sl])
        cnt[sl] += 1
    return out / cnt


def run(model, image, class_num, cfg):
    """Inferer(cfg).run(model, image, domain_label) with ``model`` closed over
    the domain label."""
    def infer(img):
        if not cfg.get("sliding_window_enable", False):
            return model(img)
        return sliding_window(model, img, class_num, cfg["sliding_window_size"], cfg["sliding_window_stride"])

    tta = cfg.get("tta_mode", 0)
    if tta == 0:
        return infer(image)
    if tta != 1:
        raise ValueError("Undefined tta_mode {0:}".format(tta))
    o1 = infer(image)
    o2 = torch.flip(infer(torch.flip(image, [-2])), [-2])
    o3 = torch.flip(infer(torch.flip(image, [-1])), [-1])
    o4 = torch.flip(infer(torch.flip(image, [-2, -1])), [-2, -1])
    return (o1 + o2 + o3 + o4) / 4
