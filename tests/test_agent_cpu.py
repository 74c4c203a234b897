"""CPU: host logic of the agent -- .cfg parsing against the reference parser's output (golden made by
oracle/gen_golden_cfg.py from the reference's own util/parse_config.py), optimiser/scheduler factory,
round-robin sharding, image-weight map, plugin errors."""
import json
import os

import numpy as np
import pytest
import torch

from fplplus_b200 import agent as A

HERE = os.path.dirname(os.path.abspath(__file__))


def test_parse_config_matches_reference_parser(tmp_path):
    with open(os.path.join(HERE, "golden", "cfg_parse.json")) as f:
        g = json.load(f)
    p = tmp_path / "x.cfg"
    p.write_text(g["text"])
    cfg = A.synchronize_config(A.parse_config(str(p)))
    assert json.loads(json.dumps(cfg)) == g["parsed"]
    # keys are lower-cased, values type-sniffed exactly like the reference
    assert cfg["testing"]["domian_label"] == 1 and cfg["network"]["feature_chns"] == [16, 32, 64, 128, 256]
    assert cfg["dataset"]["labeltoprobability_class_num"] == 2


def test_value_sniffing_edge_cases():
    f = A.parse_value_from_string
    assert f("-5") == -5 and f("1e-4") == 1e-4 and f("0.5") == 0.5
    assert f("./data/x.csv") == "./data/x.csv"
    assert f("[1, 2.5, True, None, abc]") == [1, 2.5, True, None, "abc"]
    assert f("None") is None and f("false") is False and f("DiceLoss") == "DiceLoss"


def test_optimizer_and_scheduler_factory():
    w = [torch.nn.Parameter(torch.zeros(3))]
    opt = A.get_optimizer("Adam", w, {"learning_rate": 1e-3, "momentum": 0.9, "weight_decay": 1e-5})
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["weight_decay"] == 1e-5
    sch = A.get_lr_scheduler(opt, {"lr_scheduler": "MultiStepLR", "lr_gamma": 0.5, "lr_milestones": [2, 4], "last_iter": -1})
    lrs = []
    for _ in range(5):
        opt.step()
        sch.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert np.allclose(lrs, [1e-3, 5e-4, 5e-4, 2.5e-4, 2.5e-4])
    assert A.get_lr_scheduler(opt, {"lr_scheduler": None}) is None
    with pytest.raises(ValueError):
        A.get_optimizer("LBFGS2", w, {"learning_rate": 1.0})


def test_round_robin_sharding_covers_everything_once():
    items = list(range(11))
    for world in (1, 2, 4, 8):
        parts = [A.shard_round_robin(items, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == items
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_undefined_network_and_loss_raise_like_the_reference():
    cfg = {"dataset": {"tensor_type": "float"}, "network": {"net_type": "Nope", "class_num": 2, "num_domains": 2},
           "training": {"loss_type": "NoLoss"}, "testing": {}}
    ag = A.SegmentationAgent(cfg, "train")
    with pytest.raises(ValueError, match="Undefined network"):
        ag.create_network()
    with pytest.raises(ValueError, match="Undefined loss"):
        ag.create_loss_calculator()
    with pytest.raises(ValueError):
        A.SegmentationAgent({"dataset": {"tensor_type": "double"}, "training": {}, "network": {}}, "train")


def test_npy_dataset_folds_image_weight_like_set_weight(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.standard_normal((4, 6, 8)).astype(np.float32)
    lab = rng.integers(0, 2, (4, 6, 8)).astype(np.uint8)
    w = np.where(rng.random((4, 6, 8)) < 0.3, 0.5, 1.0)
    for n, a in (("i", img), ("l", lab), ("w", w)):
        np.save(tmp_path / (n + ".npy"), a)
    ds = A.NpyVolumeDataset([("i.npy", "l.npy", "w.npy", 0.37)], class_num=2, root_dir=str(tmp_path))
    s = ds[0]
    assert s["image"].shape == (1, 4, 6, 8) and s["label_prob"].shape == (2, 4, 6, 8)
    from oracle import fpl_filter
    np.testing.assert_array_equal(s["pixel_weight"][0].numpy(), fpl_filter.set_weight_(np.float32(0.37), w.astype(np.float32)))


def test_network_plan_on_cpu_gradient_sets_and_mc_split():
    """Host-side planning of UNet2D5_dsbn needs no GPU: which parameters receive gradients (3-D convs, the selected
    domain's BN, PReLU, the transposed convs -- or the 1x1 projections in bilinear mode; never the 2-D twins), and where
    the MC-dropout sweep splits the network (first encoder level with an active dropout)."""
    import torch
    from fplplus_b200.net import UNet2D5_dsbn
    base = {"in_chns": 1, "feature_chns": [16, 32, 64, 128, 256], "dropout": [0.0, 0.0, 0.3, 0.4, 0.5],
            "conv_dims": [3, 3, 3, 3, 3], "class_num": 2, "bilinear": False, "num_domains": 2}
    net = UNet2D5_dsbn(dict(base))
    names = {id(p): n for n, p in net.named_parameters()}
    got = [names[id(p)] for p in net._grad_params(1)]
    assert len(got) == len(set(got)) == 2 + 4 * 12 + 5 * 10
    assert got[:2] == ["out_conv.weight", "out_conv.bias"] and got[-1] == "block0.conv.relu_1.weight"
    assert not any("2d" in n or ".bns.0." in n or n.endswith(".conv3d.weight") and n.count(".") == 2 for n in got)
    assert "up4.trans3d.weight" in got and sum(p.numel() for p in net._grad_params(1)) == 5648148
    bil = UNet2D5_dsbn(dict(base, bilinear=True))
    names = {id(p): n for n, p in bil.named_parameters()}
    got = [names[id(p)] for p in bil._grad_params(0)]
    assert "up4.conv3d.weight" in got and not any("trans" in n for n in got)
    # 2.5-D: the 2-D members of the first two levels take the gradients
    d25 = UNet2D5_dsbn(dict(base, conv_dims=[2, 2, 3, 3, 3]))
    names = {id(p): n for n, p in d25.named_parameters()}
    got = [names[id(p)] for p in d25._grad_params(1)]
    assert "block0.conv.conv2d_1.weight" in got and "block0.conv.conv3d_1.weight" not in got and "up4.trans2d.weight" in got
    # MC split: dropout modules in train mode with p > 0 from level 2 on
    net.eval()
    assert net._first_dropout_level() == 5
    for m in net.modules():
        if type(m) == torch.nn.Dropout:
            m.train()
    assert net._first_dropout_level() == 2
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 16, 32, 32), domain_label=torch.zeros(1, dtype=torch.long))      # CPU tensor: no fallback


def test_training_all_copies_the_graphs_static_dice_output():
    """in CUDA-graph mode train_step returns the graph's STATIC output tensors; training_all must not
    alias them (the first step of a round used to be replaced by the second).  Host logic only: a fake train_step
    that rewrites one static buffer per call, like a replay does."""
    cfg = {"dataset": {"tensor_type": "float"}, "network": {"num_domains": 2, "class_num": 2},
           "training": {"iter_valid": 3}, "testing": {}}
    ag = A.SegmentationAgent(cfg, "train")
    ag.device = torch.device("cpu")

    class _Net(object):
        def train(self):
            pass
    ag.net = _Net()
    ag.train_loaders = [[{"k": 0}], [{"k": 1}]]
    static_loss = torch.zeros(())
    static_dice = [torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)]
    calls = []

    def fake_train_step(batches):
        i = len(calls)
        calls.append(batches)
        static_loss.fill_(1.0 + i)
        static_dice[0].copy_(torch.tensor([0.1 * (i + 1), 0.2 * (i + 1)], dtype=torch.float64))
        static_dice[1].copy_(torch.tensor([0.3 * (i + 1), 0.0], dtype=torch.float64))
        return static_loss, static_dice
    ag.train_step = fake_train_step
    out = ag.training_all()
    assert len(calls) == 3
    assert abs(out["loss"] - (1 + 2 + 3) / 3 / 2) < 1e-6
    # mean over the 3 steps, then over the 2 domains
    expect = (np.array([0.2, 0.4]) + np.array([0.6, 0.0])) / 2
    np.testing.assert_allclose(out["class_dice"], expect, rtol=1e-12)


def test_data_parallel_mode_selection(monkeypatch):
    """enable_data_parallel: 'deferred' (default) = one all-reduce of the master gradient buffer per step, no hooks inside
    backward; 'overlapped' = the bucketed hook protocol; [training] grad_allreduce and $FPL_GRAD_ALLREDUCE select it."""
    class _Net(torch.nn.Module):
        grad_ready_hook = grad_wait_hook = "stale"

    cfg = {"dataset": {"tensor_type": "float"}, "network": {}, "training": {}, "testing": {}}
    monkeypatch.delenv("FPL_GRAD_ALLREDUCE", raising=False)
    ag = A.SegmentationAgent(cfg, "train")
    ag.net = _Net()
    red = ag.enable_data_parallel()
    assert isinstance(red, A.DeferredGradAllReducer) and ag.reducer is red
    assert ag.net.grad_ready_hook is None and ag.net.grad_wait_hook is None
    red = ag.enable_data_parallel("overlapped")
    assert isinstance(red, A.GradAllReducer) and ag.net.grad_ready_hook == red.hook and ag.net.grad_wait_hook == red.finish
    ag.config["training"]["grad_allreduce"] = "overlapped"
    assert isinstance(ag.enable_data_parallel(), A.GradAllReducer)
    monkeypatch.setenv("FPL_GRAD_ALLREDUCE", "deferred")
    assert isinstance(ag.enable_data_parallel(), A.DeferredGradAllReducer)
    with pytest.raises(ValueError):
        ag.enable_data_parallel("ring")
    # nothing to reduce: a step without gradients is a no-op (no process group needed)
    A.DeferredGradAllReducer(_Net()).finish()
