"""GPU timing of the CUDA-core head kernels (csrc/head.cu) at the benchmark shape.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops

DEV = "cuda:0"
L = lib.load()
p = ops.ptr


def timeit(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); fn()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for cin, k, (n, d, h, w) in [(16, 2, (4, 32, 128, 128)), (16, 5, (2, 96, 160, 160)), (32, 2, (4, 32, 128, 128))]:
    x = torch.randn((n, d, cin // 8, h, w, 8), device=DEV).to(torch.bfloat16)
    wt = torch.randn(k, cin, 1, 3, 3, device=DEV) * 0.1
    b = torch.zeros(k, device=DEV)
    logits = torch.empty((n, k, d, h, w), device=DEV)
    dl = torch.randn((n, k, d, h, w), device=DEV)
    g = torch.empty_like(x)
    dl8 = torch.empty((n, d, 1, h, w, 8), device=DEV, dtype=torch.bfloat16)
    db = torch.zeros(k, device=DEV)
    st = ops.stream_ptr()
    vox = n * d * h * w
    tf = timeit(lambda: ops.call("fpl_head_fwd", p(x), cin // 8, 0, p(wt), p(b), p(logits), n, d, h, w, cin, k, st))
    tb = timeit(lambda: ops.call("fpl_head_dgrad", p(dl), p(wt), p(g), cin // 8, 0, p(dl8), 1, 0, p(db), n, d, h, w, cin, k, st))
    bf, bb = vox * (2 * cin + 4 * k), vox * (4 * k + 2 * cin + 16)
    print("cin %d classes %d %s: fwd %.1f us (%.0f GB/s)  dgrad %.1f us (%.0f GB/s)" % (
        cin, k, "x".join(map(str, (n, d, h, w))), tf, bf / tf / 1e3, tb, bb / tb / 1e3), flush=True)
