"""Thin tensor-level wrappers over the C ABI (include/fplplus_b200.h).

Everything here takes CUDA torch tensors, hands raw pointers + the current stream to the
library and returns.  No arithmetic happens in Python/torch.
"""
import ctypes

import torch

from . import lib as _lib

call = _lib.call


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


class C8(object):
    """A C8-planar bf16 activation: ``buf`` is [N, D, c8_total, H, W, 8] bf16 and this view covers
    ``channels`` channels starting at channel group ``c8_off``."""
    __slots__ = ("buf", "c8_off", "channels")

    def __init__(self, buf, c8_off=0, channels=None):
        self.buf = buf
        self.c8_off = c8_off
        self.channels = buf.shape[2] * 8 - c8_off * 8 if channels is None else channels

    @property
    def c8_total(self):
        return self.buf.shape[2]

    @property
    def geom(self):
        n, d, _c8, h, w, _ = self.buf.shape
        return n, d, h, w

    def args(self):
        return ptr(self.buf), self.c8_total, self.c8_off


def c8_empty(n, d, c, h, w, device):
    assert c % 8 == 0
    return torch.empty((n, d, c // 8, h, w, 8), dtype=torch.bfloat16, device=device)


def ncdhw_to_c8(x):
    """fp32/bf16 [N,C,D,H,W] -> C8-planar bf16 (test/helper path; uses torch ops)."""
    n, c, d, h, w = x.shape
    pad = (-c) % 8
    if pad:
        x = torch.cat([x, x.new_zeros((n, pad, d, h, w))], 1)
    c8 = (c + pad) // 8
    return x.reshape(n, c8, 8, d, h, w).permute(0, 3, 1, 4, 5, 2).contiguous().to(torch.bfloat16)


def c8_to_ncdhw(buf, channels=None):
    n, d, c8, h, w, _ = buf.shape
    x = buf.permute(0, 2, 5, 1, 3, 4).reshape(n, c8 * 8, d, h, w).float()
    return x if channels is None else x[:, :channels]


def is_sm100():
    return bool(_lib.load().fpl_device_is_sm100())
