"""GPU probe: clock64() trace of CTA 0 of the depth-folded conv kernel (fpl_debug_set 47): where each role waits.
roles: 0 producer (empty slot obtained), 1 MMA warp (box landed), 2 MMA warp (box issued + committed),
       3 epilogue warp 2 (output plane complete), 4 MMA warp (accumulators free: item start).  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fplplus_b200 import lib, ops
from tests._util import bf16_round, randn, to_c8

DEV = "cuda:0"
L = lib.load()
P = ops.ptr


def run(cin, cout, shape, knobs, label):
    n, d, h, w = shape
    x = to_c8(bf16_round(randn(1, n, cin, d, h, w)).to(DEV))
    wt = bf16_round(randn(2, cout, cin, 3, 3, 3, scale=0.1)).to(DEV)
    y = torch.zeros((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    img = torch.empty(L.fpl_conv3d_dfold_image_bytes(cin, cout) // 2, dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_conv3d_dfold_prep_weight", P(wt), cin, cout, 0, P(img), ops.stream_ptr())
    trace = torch.zeros(7 * 4096, dtype=torch.int64, device=DEV)
    for k in (40, 41, 42, 43, 44, 45, 46):
        L.fpl_debug_set(k, 0)
    for k, v in knobs.items():
        L.fpl_debug_set(k, v)
    fn = lambda: ops.call("fpl_conv3d_tc_dfold", P(x), cin // 8, 0, P(img), None, P(y), cout // 8, 0, None, n, d, h, w, cin, cout, ops.stream_ptr())
    fn(); fn()
    torch.cuda.synchronize()
    L.fpl_debug_set(47, trace.data_ptr())
    fn()
    torch.cuda.synchronize()
    L.fpl_debug_set(47, 0)
    t = trace.cpu().numpy().reshape(7, 4096)
    t0 = min(int(r[r > 0].min()) for r in t if (r > 0).any())
    print("== %s  %d->%d %s knobs %s" % (label, cin, cout, shape, knobs))
    names = ["producer slot", "mma box landed", "mma box issued", "epi plane done", "mma item start", "producer issued", "box landed (watcher)"]
    for role in range(7):
        r = t[role][t[role] > 0] - t0
        if len(r) == 0:
            continue
        dd = np.diff(r)
        print("  %-15s n=%3d first %6d last %7d | mean step %6.0f median %6.0f max %6d | first 24 steps: %s" % (
            names[role], len(r), r[0], r[-1], dd.mean() if len(dd) else 0, np.median(dd) if len(dd) else 0, dd.max() if len(dd) else 0,
            " ".join(str(int(v)) for v in dd[:24])))
    iss, land, slot, done = t[5][t[5] > 0] - t0, t[6][t[6] > 0] - t0, t[0][t[0] > 0] - t0, t[2][t[2] > 0] - t0
    m = min(len(iss), len(land))
    lat = land[:m] - iss[:m]
    print("  TMA latency issue->landed: mean %.0f median %.0f min %d max %d | first 24: %s" % (lat.mean(), np.median(lat), lat.min(), lat.max(), " ".join(str(int(v)) for v in lat[:24])))
    print("  producer issue cost (slot->issued): mean %.0f" % (iss[:len(slot)] - slot[:len(iss)]).mean())
    st = knobs.get(43, 0)
    for k in (40, 41, 42, 43, 44, 45, 46):
        L.fpl_debug_set(k, 0)


LO = {44: 1, 42: 2}
for knobs, label in [({46: 1}, "base pb1"), ({46: 2}, "base pb2"), ({**LO, 46: 1}, "load-only pb1"), ({**LO, 46: 2}, "load-only pb2"),
                     ({**LO, 46: 1, 40: 1}, "load-only pb1 1cta"), ({42: 2, 46: 2}, "noepi pb2"), ({44: 1, 46: 2}, "nomma pb2")]:
    run(16, 16, (4, 32, 128, 128), knobs, label)
run(32, 16, (4, 32, 128, 128), {46: 2}, "base pb2")
run(32, 16, (4, 32, 128, 128), {**LO, 46: 2}, "load-only pb2")
