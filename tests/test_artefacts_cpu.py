"""CPU: artefact formats of the FPL+ recipe -- NIfTI-1 I/O without SimpleITK, the training CSV, the
image-weight table (the reference's missing `get image_weight.py`) against the shipped artefacts, pixel-weight volumes,
Dice / ASSD evaluation."""
import csv
import gzip
import json
import os
import struct

import numpy as np
import pytest

from fplplus_b200 import artefacts as A

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("ext", [".nii.gz", ".nii"])
@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
def test_nifti_round_trip_keeps_voxels_and_geometry(tmp_path, ext, dtype):
    rng = np.random.default_rng(3)
    data = (rng.standard_normal((5, 7, 9)) * 50).astype(dtype)
    spacing = (0.41, 0.43, 1.5)
    origin = (-101.5, 77.25, 12.0)
    th = 0.3                                                     # an oblique direction matrix (rotation about z) in LPS
    direction = (np.cos(th), -np.sin(th), 0.0, np.sin(th), np.cos(th), 0.0, 0.0, 0.0, 1.0)
    p = str(tmp_path / ("vol" + ext))
    A.write_nifti(data, p, spacing, origin, direction)
    img = A.read_nifti(p)
    assert img["data"].dtype == np.dtype(dtype) and img["data"].shape == (5, 7, 9)
    np.testing.assert_array_equal(img["data"], data)
    np.testing.assert_allclose(img["spacing"], spacing, rtol=1e-6)
    np.testing.assert_allclose(img["origin"], origin, rtol=1e-6)
    np.testing.assert_allclose(img["direction"], direction, atol=1e-6)
    # the PyMIC-shaped loader: [C, D, H, W], spacing reordered to (z, y, x)
    d = A.load_nifty_volume_as_4d_array(p)
    assert d["data_array"].shape == (1, 5, 7, 9)
    np.testing.assert_allclose(d["spacing"], spacing[::-1], rtol=1e-6)
    # header fields a NIfTI reader checks
    raw = (gzip.open(p) if ext.endswith("gz") else open(p, "rb")).read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\0"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 9, 7, 5)          # x fastest
    assert struct.unpack("<f", raw[108:112])[0] == 352.0


def test_nifti_reader_handles_qform_only_big_endian_and_scaling(tmp_path):
    """A hand-built header: big endian, qform only (identity rotation, RAS), scl_slope/inter, int16."""
    vox = np.arange(2 * 3 * 4, dtype=">i2").reshape(2, 3, 4)
    hdr = bytearray(348)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, 4, 3, 2, 1, 1, 1, 1)
    struct.pack_into(">h", hdr, 70, 4)
    struct.pack_into(">h", hdr, 72, 16)
    struct.pack_into(">8f", hdr, 76, 1.0, 0.5, 0.6, 2.0, 0, 0, 0, 0)
    struct.pack_into(">f", hdr, 108, 352.0)
    struct.pack_into(">2f", hdr, 112, 2.0, 10.0)
    struct.pack_into(">2h", hdr, 252, 1, 0)
    struct.pack_into(">3f", hdr, 256, 0.0, 0.0, 0.0)
    struct.pack_into(">3f", hdr, 268, 5.0, 6.0, 7.0)
    hdr[344:348] = b"n+1\0"
    p = tmp_path / "be.nii"
    p.write_bytes(bytes(hdr) + b"\0" * 4 + vox.tobytes())
    img = A.read_nifti(str(p))
    np.testing.assert_allclose(img["data"], vox.astype(np.float64) * 2.0 + 10.0)
    assert img["spacing"] == (0.5, 0.6000000238418579, 2.0)
    np.testing.assert_allclose(img["origin"], (-5.0, -6.0, 7.0))                      # RAS -> LPS
    np.testing.assert_allclose(img["direction"], (-1, 0, 0, 0, -1, 0, 0, 0, 1), atol=1e-7)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.nii"
        bad.write_bytes(b"\0" * 400)
        A.read_nifti(str(bad))


def test_save_with_reference_geometry_and_pixel_weight_files(tmp_path):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2, (4, 6, 8)).astype(np.uint8)
    b = rng.integers(0, 2, (4, 6, 8)).astype(np.uint8)
    ref = str(tmp_path / "t.nii.gz")
    A.write_nifti(a, ref, (0.5, 0.5, 1.5), (1.0, 2.0, 3.0))
    other = str(tmp_path / "s.nii.gz")
    A.save_nd_array_as_image(b, other, ref)                                   # geometry copied from the reference image
    assert A.read_nifti(other)["spacing"] == A.read_nifti(ref)["spacing"]
    w = A.pixel_weight_from_label_files(ref, other, str(tmp_path / "w.nii.gz"))
    from oracle import fpl_filter
    np.testing.assert_array_equal(w, fpl_filter.agreement_weight(a, b))        # data/get_pixel_weight.py:21-26
    back = A.read_nifti(str(tmp_path / "w.nii.gz"))
    np.testing.assert_array_equal(back["data"], w)
    np.testing.assert_allclose(back["origin"], (1.0, 2.0, 3.0))


def test_image_weight_table_and_csv_reproduce_the_shipped_artefacts(tmp_path):
    """PRODUCT functions (fpl.image_weights, artefacts.image_weight_table / train_csv_from_uncertainty) against the only
    golden artefacts the reference ships: dataset/weight/cyc121_vst1s-gan.npy -> config_dual/data_vs/train_vs_t1s_wi+wp.csv."""
    from fplplus_b200 import fpl
    with open(os.path.join(HERE, "golden", "fpl_image_weights.json")) as f:
        g = json.load(f)
    u = [1 if s else v for v, s in zip(g["uncertainty"], g["sentinel"])]
    np.testing.assert_allclose(fpl.image_weights(u), g["csv_image_weight"], rtol=0, atol=1e-12)
    srt = [([v], n) for v, n in zip(u, g["names"])]                          # the object-array rows of agent_seg.py:957-960
    table = A.image_weight_table(srt)
    assert [n for n, _w in table] == g["csv_names"]
    np.testing.assert_allclose([w for _n, w in table], g["csv_image_weight"], rtol=0, atol=1e-12)
    # shuffled input through the product sort gives the shipped order again
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(u))
    again = fpl.sort_uncertainty({g["names"][i]: [u[i]] for i in perm})
    assert [n for _v, n in again] == g["names"]
    out_csv = str(tmp_path / "train.csv")
    rows = A.train_csv_from_uncertainty(again, out_csv, label_of=lambda n: "lab/" + n, pixel_weight_of=lambda n: "pw/" + n)
    assert len(rows) == len(u)
    with open(out_csv) as f:
        rd = list(csv.reader(f))
    assert rd[0] == ["image", "label", "pixel_weight", "image_weight"]            # train_vs_t1s_wi+wp.csv header
    assert rd[1][0] == g["csv_names"][0] and rd[1][1] == "lab/" + g["csv_names"][0] and rd[1][2] == "pw/" + g["csv_names"][0]
    np.testing.assert_allclose([float(r[3]) for r in rd[1:]], g["csv_image_weight"], rtol=0, atol=1e-12)
    assert float(rd[1][3]) == 1.01 and float(rd[-1][3]) == 0.01


def test_dice_and_assd_closed_forms(tmp_path):
    s = np.zeros((12, 20, 20), np.uint8)
    g = np.zeros_like(s)
    s[3:9, 5:15, 5:15] = 1
    g[3:9, 5:15, 6:16] = 1                                                       # the same box shifted by one voxel in x
    inter, vs = 6 * 10 * 9, 6 * 10 * 10
    assert abs(A.binary_dice(s, g) - (2.0 * inter + 1e-5) / (2 * vs + 1e-5)) < 1e-12
    assert A.binary_dice(s, s) == pytest.approx(1.0)
    assert A.binary_assd(s, s) == 0
    e = A.get_edge_points(s)
    assert e.sum() == vs - 4 * 8 * 8                                             # the shell of a 6x10x10 box
    a1 = A.binary_assd(s, g)
    assert 0.0 < a1 < 1.0                                                         # faces moved by 1 voxel, the rest overlap
    a2 = A.binary_assd(s, g, spacing=[1.0, 1.0, 3.0])
    assert a2 > a1                                                               # anisotropic spacing stretches x distances
    assert A.binary_assd(s, np.zeros_like(s)) == 50                              # empty mask: capped sentinel
    # folder evaluation + CSV layout of evaluation_seg_train.py:560-575
    for name, arr in (("s.nii.gz", s), ("g.nii.gz", g)):
        A.write_nifti(arr, str(tmp_path / name))
    rows = A.evaluate_folder([("case0", str(tmp_path / "s.nii.gz"), str(tmp_path / "g.nii.gz"))], [1], str(tmp_path / "dice.csv"))
    assert rows[0][0] == "case0" and rows[1][0] == "mean" and rows[2][0] == "std"
    assert abs(rows[0][1] - A.binary_dice(s, g)) < 1e-12
    with open(tmp_path / "dice.csv") as f:
        assert f.readline().strip() == "image,class_1"
