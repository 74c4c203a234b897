"""GPU: the agent end to end on synthetic loaders -- training_all rounds with the reference's checkpoint
protocol, resume, pseudo-label inference and the FPL image-uncertainty branch."""
import os

import numpy as np
import pytest
import torch

from oracle import fpl_filter, inferer as oracle_inferer, synth, unet_dsbn
from oracle.gen_golden import NET_PARAMS

pytestmark = pytest.mark.gpu
SHAPE = (16, 32, 32)


def _batches(seed, n_batches, bs, weighted):
    out = []
    for i in range(n_batches):
        lab = synth.synth_label(bs, 2, SHAPE, seed=seed + i)
        x = synth.synth_image(bs, 1, SHAPE, seed=seed + i) * 0.5 + (lab[:, None] > 0) * 2.0
        b = {"image": torch.from_numpy(x.astype(np.float32)), "label_prob": torch.from_numpy(synth.one_hot(lab, 2)),
             "names": ["case_%d_%d" % (seed, i)] * bs}
        if weighted:
            pw, iw = synth.synth_pixel_weight(lab, seed=seed + i)
            b["pixel_weight"], b["image_weight"] = torch.from_numpy(pw), torch.from_numpy(iw)
        out.append(b)
    return out


def _config(tmp_path, **train):
    tr = {"train_fpl_uda": True, "dual": True, "dis": False, "val_t1": False, "val_t2": False, "gpus": [0],
          "loss_type": ["DiceLoss", "CrossEntropyLoss"], "loss_weight": [0.5, 0.5], "optimizer": "Adam",
          "learning_rate": 2e-3, "momentum": 0.9, "weight_decay": 1e-5, "lr_scheduler": "MultiStepLR", "lr_gamma": 0.5,
          "lr_milestones": [6], "ckpt_save_dir": str(tmp_path / "model" / "vs_S"), "iter_start": 0, "iter_max": 8,
          "iter_valid": 4, "iter_save": 4, "early_stop_patience": None}
    tr.update(train)
    te = {"fpl": False, "gpus": [0], "domian_label": 1, "ae": False, "ckpt_mode": 0, "evaluation_mode": True,
          "test_time_dropout": False, "tta_mode": 1, "sliding_window_enable": True, "sliding_window_size": [16, 32, 32],
          "sliding_window_stride": [8, 16, 32], "output_dir": str(tmp_path / "result"),
          "fpl_uncertainty_sorted": str(tmp_path / "sorted.npy")}
    return {"dataset": {"tensor_type": "float", "train_batch_size": 2}, "network": dict(NET_PARAMS),
            "training": tr, "testing": te}


def test_train_valid_checkpoints_resume_and_infer(tmp_path):
    from fplplus_b200.agent import SegmentationAgent
    cfg = _config(tmp_path)
    ag = SegmentationAgent(cfg, "train")
    ag.set_loaders(train=[_batches(10, 3, 2, False), _batches(20, 3, 2, True)],
                   valid=[_batches(30, 1, 1, False), _batches(40, 1, 1, False)])
    hist = ag.run()
    assert [h[0] for h in hist] == [4, 8]
    assert hist[-1][1]["loss"] < hist[0][1]["loss"]                       # it trains
    assert 0.0 <= hist[-1][2]["avg_dice"] <= 1.0 and hist[-1][1]["class_dice"].shape == (2,)
    d = cfg["training"]["ckpt_save_dir"]
    assert open(os.path.join(d, "vs_S_latest.txt")).read() == "8"
    best_it = int(open(os.path.join(d, "vs_S_best.txt")).read())
    ck = torch.load(os.path.join(d, "vs_S_%d.pt" % best_it), map_location="cpu", weights_only=False)
    assert set(ck) == {"iteration", "valid_pred", "model_state_dict", "optimizer_state_dict"}
    assert len(ck["model_state_dict"]) == 484
    # scheduler: MultiStepLR milestone 6 halved the rate by iteration 8
    assert abs(ag.current_lr() - 1e-3) < 1e-12 and abs(float(ag.optimizer.param_groups[0]["lr"]) - 1e-3) < 1e-9

    # resume from iteration 8 (agent_seg.py:721-734) and run one more round
    cfg2 = _config(tmp_path, iter_start=8, iter_max=12, iter_save=12)
    ag2 = SegmentationAgent(cfg2, "train")
    ag2.set_loaders(train=[_batches(10, 3, 2, False), _batches(20, 3, 2, True)],
                    valid=[_batches(30, 1, 1, False), _batches(40, 1, 1, False)])
    hist2 = ag2.run()
    assert [h[0] for h in hist2] == [12]
    assert open(os.path.join(d, "vs_S_latest.txt")).read() == "12"

    # pseudo labels through the agent (ckpt_mode 0 = latest) vs the oracle on the same checkpoint
    vol = synth.synth_image(1, 1, (24, 48, 48), seed=77)
    cfg3 = _config(tmp_path)
    ag3 = SegmentationAgent(cfg3, "test")
    ag3.set_loaders(test=[{"image": torch.from_numpy(vol), "names": ["vol_a.nii.gz"]}])
    out = ag3.run()
    lab = out["vol_a.nii.gz"]
    assert lab.dtype == np.uint8 and lab.shape == (1, 24, 48, 48)
    # NIfTI name in -> NIfTI label volume out (artefacts.py, no SimpleITK), same voxels
    from fplplus_b200 import artefacts
    saved = artefacts.read_nifti(os.path.join(cfg3["testing"]["output_dir"], "vol_a.nii.gz"))
    assert saved["data"].dtype == np.uint8 and np.array_equal(saved["data"], lab[0])
    ck = torch.load(os.path.join(d, "vs_S_12.pt"), map_location="cpu", weights_only=False)
    st = {k: v.float() if v.is_floating_point() else v for k, v in ck["model_state_dict"].items()}
    params = dict(NET_PARAMS)

    def model(x):
        return unet_dsbn.forward(st, x, 1, params)
    with torch.no_grad():
        ref = oracle_inferer.run(model, torch.from_numpy(vol), 2, cfg3["testing"])
    ref_lab = fpl_filter.pseudo_label(ref.numpy())
    agree = float((ref_lab == lab).mean())
    print("agent pseudo labels vs oracle: agreement %.5f" % agree)
    assert agree >= 0.995


def test_fpl_branch_sorts_uncertainties(tmp_path):
    from fplplus_b200.agent import SegmentationAgent
    cfg = _config(tmp_path)
    cfg["testing"].update(fpl=True, test_time_dropout=True, ckpt_mode=2, tta_mode=0)
    ag = SegmentationAgent(cfg, "test")
    ag.create_network()
    sd = synth.synth_state_dict()
    ag.net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    vols = [{"image": torch.from_numpy(synth.synth_image(1, 1, (16, 32, 32), seed=80 + i)), "names": ["v%d" % i]} for i in range(3)]
    ag.set_loaders(test=vols)
    torch.manual_seed(5)
    srt = ag.infer(load_checkpoint=False)
    assert [n for _v, n in srt] == sorted([n for _v, n in srt], key=lambda n: dict((k, v) for v, k in srt)[n])
    vals = [v[0] for v, _n in srt]
    assert vals == sorted(vals) and all(v == 1 or 0 < v < 1 for v in vals)
    saved = np.load(cfg["testing"]["fpl_uncertainty_sorted"], allow_pickle=True)
    assert saved.shape == (3, 2)                                          # the reference's object-array layout
    # MC dropout was live (K passes differ) -> a non-sentinel uncertainty somewhere on an untrained net
    assert any(v != 1 for v in vals)


def test_validation_loss_and_dice_match_the_oracle(tmp_path):
    """a19 (agent_seg.py:509-604): per domain, Inferer pass over every validation volume, unweighted loss of the batch and
    per-volume hard Dice; model selection by val_t1 / val_t2.  Against the oracle Inferer + losses on the same weights."""
    from oracle import losses
    from fplplus_b200.agent import SegmentationAgent
    cfg = _config(tmp_path)
    ag = SegmentationAgent(cfg, "train")
    ag.create_network()
    sd = synth.synth_state_dict()
    ag.net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    ag._pick_device("training")
    ag.net.to(ag.device)
    ag.create_loss_calculator()
    shape = (24, 48, 48)

    def vols(seed, n):
        out = []
        for i in range(n):
            lab = synth.synth_label(1, 2, shape, seed=seed + i)
            x = synth.synth_image(1, 1, shape, seed=seed + i) * 0.5 + (lab[:, None] > 0) * 2.0
            out.append({"image": torch.from_numpy(x.astype(np.float32)), "label_prob": torch.from_numpy(synth.one_hot(lab, 2)),
                        "names": ["v%d" % (seed + i)]})
        return out
    valid = [vols(300, 2), vols(400, 3)]
    ag.set_loaders(valid=valid)
    st = unet_dsbn.to_torch_state(sd, requires_grad=False)
    params = dict(NET_PARAMS)
    ref = []
    for d in (0, 1):
        ls, ds = [], []
        for b in valid[d]:
            with torch.no_grad():
                z = oracle_inferer.run(lambda x: unet_dsbn.forward(st, x, d, params), b["image"], 2, cfg["testing"])
            ls.append(float(losses.combined_loss(z, b["label_prob"], None, 0.5, 0.5)))
            ds.append(losses.hard_dice(z, b["label_prob"]).numpy())
        ref.append((np.mean(ls), np.mean(np.stack(ds, 0), 0)))
    for key, pick in ((None, (0, 1)), ("val_t1", (0,)), ("val_t2", (1,))):
        cfg["training"]["val_t1"], cfg["training"]["val_t2"] = key == "val_t1", key == "val_t2"
        got = ag.validation()
        exp_loss = float(np.mean([ref[d][0] for d in pick]))
        exp_dice = np.mean(np.stack([ref[d][1] for d in pick], 0), 0)
        print(key, "validation loss %.5f oracle %.5f; class dice %s oracle %s" % (got["loss"], exp_loss, got["class_dice"], exp_dice))
        assert abs(got["loss"] - exp_loss) <= 1e-2 * abs(exp_loss)
        np.testing.assert_allclose(got["class_dice"], exp_dice, atol=2e-2)
        assert abs(got["avg_dice"] - float(exp_dice.mean())) <= 2e-2
    assert ag.net.training                                    # validation() hands the network back in train mode
