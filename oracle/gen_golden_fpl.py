"""Golden vectors for the FPL branch of the reference agent's ``infer()``.

``pymic.net_run_dsbn.agent_seg`` cannot be imported as shipped (tensorboardX,
SimpleITK, GeodisTK, scikit-image and the whole ``pymic.net.net2d`` package are
missing), so the missing modules are replaced by inert stubs and the REAL
``SegmentationAgent.infer`` (agent_seg.py:834-964) is driven with a fake
loader / network and an inferer that replays prepared logits.  What the
reference then computes -- softmax, per-voxel variance, class-1 mean, entropy
map, boundary count, the ``<50 => 1`` sentinel, the ascending (value, name)
sort and the object-array ``.npy`` -- is exactly its own code.

Build container only:  python -m oracle.gen_golden_fpl
"""
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
K_PASSES = 6


def fpl_case_logits(seed, shape=(12, 24, 24), confident=False):
    """K MC passes [1,2,D,H,W] fp32: a blob of class-1 evidence plus per-pass noise."""
    g = np.random.Generator(np.random.PCG64(1000 + seed))
    D, H, W = shape
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing="ij")
    c = [g.uniform(0.3, 0.7) * s for s in shape]
    r = [g.uniform(0.15, 0.3) * s for s in shape]
    dist = np.sqrt(((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2)
    base = (1.0 - dist) * 4.0
    passes = []
    for _ in range(K_PASSES):
        z = np.zeros((1, 2) + shape, np.float32)
        if confident:
            z[0, 0] = 12.0 + 0.01 * g.standard_normal(shape)
            z[0, 1] = -12.0
        else:
            z[0, 1] = base + g.standard_normal(shape) * 0.7
            z[0, 0] = -z[0, 1] * 0.5 + g.standard_normal(shape) * 0.3
        passes.append(z)
    return passes


CASES = [("vs_gk_7_t2.nii.gz", 7, False), ("vs_gk_12_t2.nii.gz", 12, False),
         ("vs_gk_93_t2.nii.gz", 93, True), ("vs_gk_3_t2.nii.gz", 3, False),
         ("vs_gk_111_t2.nii.gz", 111, True), ("vs_gk_40_t2.nii.gz", 40, False)]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return mock.MagicMock(name=f"{self.__name__}.{name}")


def _install_stubs():
    names = ["tensorboardX", "SimpleITK", "GeodisTK", "skimage", "skimage.measure", "skimage.morphology",
             "skimage.transform", "skimage.filters", "matplotlib", "matplotlib.pyplot", "pymic.net.net2d"]
    names += ["pymic.net.net2d." + m for m in
              ("unet2d", "unet2d_dual_branch", "unet2d_urpc", "unet2d_cct", "cople_net", "unet2d_attention",
               "unet2d_nest", "unet2d_scse", "unet2d_multi_decoder", "unet2d_canet")]
    for n in names:
        try:
            __import__(n)
        except Exception:
            sys.modules[n] = _Stub(n)


class _FakeImage:
    def __init__(self, shape):
        self.shape = shape

    def float(self):
        return self

    def double(self):
        return self

    def to(self, *_a, **_k):
        return self


class _ReplayInferer:
    def __init__(self, table):
        self.table, self.idx, self.current = table, 0, None

    def run(self, net, images, domain_label):
        name = self.current
        out = torch.from_numpy(self.table[name][self.idx % K_PASSES])
        self.idx += 1
        return out


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "PyMIC"))
    _install_stubs()
    from pymic.net_run_dsbn.agent_seg import SegmentationAgent

    table = {n: fpl_case_logits(s, confident=c) for n, s, c in CASES}
    inferer = _ReplayInferer(table)

    class _Loader:
        def __iter__(self):
            for n, _s, _c in CASES:
                inferer.current, inferer.idx = n, 0
                yield {"image": _FakeImage((1, 1, 12, 24, 24)), "names": [n]}

    tmp = tempfile.mkdtemp()
    out_npy = os.path.join(tmp, "sorted.npy")
    agent = object.__new__(SegmentationAgent)
    agent.config = {"testing": {"domian_label": 1, "gpus": [0], "fpl": True, "ae": None, "ckpt_mode": 2,
                                "ckpt_name": os.path.join(tmp, "x.pt"), "evaluation_mode": True,
                                "test_time_dropout": True, "fpl_uncertainty_sorted": out_npy},
                    "network": {"class_num": 2}, "training": {}, "dataset": {}}
    agent.net = mock.MagicMock()
    agent.inferer = inferer
    agent.postprocessor = None
    agent.transform_list = []
    agent.test_loader = _Loader()
    agent.tensor_type = "float"
    saved = {}

    def _capture(path, obj):  # numpy>=1.24 refuses the ragged list the reference hands to np.save
        saved["path"], saved["obj"] = path, obj

    with mock.patch.object(torch, "load", return_value={"model_state_dict": {}}), \
            mock.patch.object(np, "save", _capture):
        agent.infer()
    assert saved["path"] == out_npy
    arr = saved["obj"]
    names = [str(r[1]) for r in arr]
    values = np.asarray([float(r[0][0]) for r in arr], np.float64)
    is_sentinel = np.asarray([isinstance(r[0][0], int) for r in arr])
    np.savez_compressed(os.path.join(GOLD, "fpl_infer.npz"), names=np.asarray(names), values=values,
                        is_sentinel=is_sentinel,
                        case_names=np.asarray([c[0] for c in CASES]), case_seeds=np.asarray([c[1] for c in CASES]),
                        case_confident=np.asarray([c[2] for c in CASES]))
    print("fpl_infer", list(zip(names, values)))


if __name__ == "__main__":
    main()
