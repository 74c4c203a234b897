"""GPU probe: where does the depth-folded conv kernel spend its time?  Graph-timed launches under the timing knobs
(fpl_debug_set 40..44: CTAs per SM, planes per depth chunk, epilogue without stores / without TMEM reads, stage count,
no MMAs).  Development tool; the knobs produce wrong results by design."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fplplus_b200 import lib, ops
from tests._util import bf16_round, randn, to_c8
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_util import graph_time

DEV = "cuda:0"
L = lib.load()
P = ops.ptr

LOADONLY = {44: 1, 42: 2}
VARIANTS = [("base", {}), ("noepi", {42: 2}), ("load-only", LOADONLY), ("lo noWhalo", {**LOADONLY, 45: 1}),
            ("lo u64", {**LOADONLY, 45: 2}), ("lo L2-256", {**LOADONLY, 45: 4}), ("lo noHhalo", {**LOADONLY, 45: 8}),
            ("lo no halo", {**LOADONLY, 45: 9}), ("lo 1cta", {**LOADONLY, 40: 1}), ("noWhalo", {45: 1}), ("u64", {45: 2}),
            ("L2-256", {45: 4}), ("no halo", {45: 9})]
if len(sys.argv) > 1 and sys.argv[1] == "stages":
    VARIANTS = [("lo s%d" % k, {**LOADONLY, 43: k}) for k in (2, 3, 4, 6, 8)] + [("lo 1cta s%d" % k, {**LOADONLY, 40: 1, 43: k}) for k in (2, 4, 8)] + \
               [("noepi s%d" % k, {42: 2, 43: k}) for k in (2, 4, 8)] + [("nomma s%d" % k, {44: 1, 43: k}) for k in (2, 4, 8)]
if len(sys.argv) > 1 and sys.argv[1] == "pb":
    VARIANTS = [("pb%d" % k, {46: k}) for k in (1, 2, 3, 6)] + [("lo pb%d" % k, {**LOADONLY, 46: k}) for k in (1, 2, 3, 6)] + \
               [("pb2 s%d" % k, {46: 2, 43: k}) for k in (2, 3, 4)] + [("pb3 s%d" % k, {46: 3, 43: k}) for k in (2, 3)] + [("noepi pb2", {42: 2, 46: 2})]
if len(sys.argv) > 1 and sys.argv[1] == "tail":
    VARIANTS = [("split tail", {}), ("no split", {48: 0}), ("pb2", {46: 2}), ("pb2 no split", {46: 2, 48: 0}), ("pb1 no split", {46: 1, 48: 0})]
if len(sys.argv) > 1 and sys.argv[1] == "round1":
    VARIANTS = [("base", {}), ("1cta", {40: 1}), ("dc8", {41: 8}), ("nostore", {42: 1}), ("noepi", {42: 2}), ("nomma", {44: 1}),
                ("nomma+noepi", {44: 1, 42: 2}), ("stages8", {43: 8}), ("stages3", {43: 3}), ("1cta+noepi", {40: 1, 42: 2})]


def probe(cin, cout, shape, with_stats):
    n, d, h, w = shape
    x = to_c8(bf16_round(randn(1, n, cin, d, h, w)).to(DEV))
    wt = bf16_round(randn(2, cout, cin, 3, 3, 3, scale=0.1)).to(DEV)
    y = torch.zeros((n, d, cout // 8, h, w, 8), dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    img = torch.empty(L.fpl_conv3d_dfold_image_bytes(cin, cout) // 2, dtype=torch.bfloat16, device=DEV)
    ops.call("fpl_conv3d_dfold_prep_weight", P(wt), cin, cout, 0, P(img), ops.stream_ptr())
    fn = lambda: ops.call("fpl_conv3d_tc_dfold", P(x), cin // 8, 0, P(img), None, P(y), cout // 8, 0, P(stats) if with_stats else None,
                          n, d, h, w, cin, cout, ops.stream_ptr())
    row = []
    for name, knobs in VARIANTS:
        for k in (40, 41, 42, 43, 44, 45, 46):
            L.fpl_debug_set(k, 0)
        L.fpl_debug_set(48, 1)
        for k, v in knobs.items():
            L.fpl_debug_set(k, v)
        try:
            row.append("%s %.1f" % (name, graph_time(fn)))
        except Exception:
            row.append("%s n/a" % name)
    for k in (40, 41, 42, 43, 44, 45, 46):
        L.fpl_debug_set(k, 0)
    print("%d->%d %s stats=%d | " % (cin, cout, shape, with_stats) + " | ".join(row), flush=True)


for args in [(16, 16, (4, 32, 128, 128), 0), (16, 16, (4, 32, 128, 128), 1), (32, 16, (4, 32, 128, 128), 0), (16, 32, (4, 16, 64, 64), 0),
             (32, 32, (4, 16, 64, 64), 0), (64, 32, (4, 16, 64, 64), 0), (16, 32, (4, 32, 128, 128), 0)]:
    probe(*args)
