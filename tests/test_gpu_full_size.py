"""GPU: the BASELINE.json configurations that are not the bench workload, at their FULL named sizes.

  configs[3]  BraTS-style 2-class DSBN U-Net, 1x128x128x128 patches
  configs[4]  MMWHS 5-class DSBN U-Net, 1x96x160x160 patches
  configs[1]  48x256x256 volumes through the filter kernels

The oracle is only asked for what it finishes in seconds at these sizes (one fp32 eval forward of one sample on the
host cores); the rest is checked through size-independent properties: linearity of backward in the loss scale,
gradients only where the reference has them, bit-exact labels / agreement weights against numpy, window stitching
identity."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle.gen_golden import NET_PARAMS
from tests._util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CFG3 = (dict(NET_PARAMS, class_num=2, dropout=[0.0, 0.0, 0.3, 0.4, 0.5]), (128, 128, 128))
CFG4 = (dict(NET_PARAMS, class_num=5, dropout=[0.0, 0.0, 0.3, 0.4, 0.5]), (96, 160, 160))


def _net(params, seed=1):
    from fplplus_b200.net import UNet2D5_dsbn
    net = UNet2D5_dsbn(dict(params))
    sd = synth.synth_state_dict(params["in_chns"], params["feature_chns"], params["class_num"], params["num_domains"], seed=seed)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return net.to(DEV), sd


@pytest.mark.parametrize("cfg", [CFG3, CFG4], ids=["brats_128cube", "mmwhs_96x160x160_5class"])
def test_full_size_eval_forward_matches_oracle(cfg):
    """One sample, eval mode (running statistics), against the fp32 CPU oracle: logits rel 1e-2, argmax labels >= 99.9 %
    on the voxels the tolerance can decide (the untrained synthetic net crowds margins around zero)."""
    from oracle import unet_dsbn
    params, shape = cfg
    net, sd = _net(params)
    net.eval()
    x = torch.from_numpy(synth.synth_image(1, 1, shape, seed=21))
    with torch.no_grad():
        out = net(x.to(DEV), domain_label=torch.zeros(1, dtype=torch.long)).cpu()
        ref = unet_dsbn.forward(unet_dsbn.to_torch_state(sd), x, 0, params, bn_training=False)
    assert out.shape == (1, params["class_num"]) + shape
    err = rel_l2(out, ref)
    top2 = ref.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-2 * ref.abs().max()
    agree = float((out.argmax(1) == ref.argmax(1))[decided].float().mean())
    print("full-size eval logits rel_l2 %.2e, decided-voxel label agreement %.5f (%.2f of the volume)" %
          (err, agree, float(decided.float().mean())))
    assert err < 1e-2
    assert agree >= 0.999


@pytest.mark.parametrize("cfg", [CFG3, CFG4], ids=["brats_128cube", "mmwhs_96x160x160_5class"])
def test_full_size_train_step_properties(cfg):
    """Batch 2 train-mode forward + weighted Dice+CE + backward at the named patch size.  Properties: finite loss and
    gradients; exactly the reference's set of parameters receives gradients (3-D convs, the selected domain's BN, PReLU;
    never the 2-D twins, 1x1 convs or the other domain's BN); backward is linear in the loss scale; running statistics
    of the selected domain moved, the other domain's did not."""
    from fplplus_b200.loss import CombinedLoss
    from fplplus_b200.registry import loss_dict
    params, shape = cfg
    c = params["class_num"]
    net, sd = _net(params)
    net.train()
    crit = CombinedLoss({"loss_type": ["DiceLoss", "CrossEntropyLoss"], "loss_weight": [0.5, 0.5]}, loss_dict)
    x = torch.from_numpy(synth.synth_image(2, 1, shape, seed=31)).to(DEV)
    lab = synth.synth_label(2, c, shape, seed=31)
    y = torch.from_numpy(synth.one_hot(lab, c)).to(DEV)
    pw = torch.from_numpy(synth.synth_pixel_weight(lab, seed=31)[0]).to(DEV)
    dom = torch.ones(2, dtype=torch.long)

    def grads(scale):
        for p in net.parameters():
            p.grad = None
        torch.manual_seed(3)                     # same dropout stream
        out = net(x, domain_label=dom)
        loss = crit({"prediction": out, "ground_truth": y, "pixel_weight": pw})
        (loss * scale).backward()
        torch.cuda.synchronize()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}

    rm_before = {k: v.clone() for k, v in net.state_dict().items() if "running_mean" in k}
    loss1, g1 = grads(1.0)
    assert np.isfinite(loss1) and 0.0 < loss1 < 5.0
    for k, g in g1.items():
        assert torch.isfinite(g).all(), k
        assert "2d" not in k and ".bns.0." not in k, k
        assert not (k.startswith("up") and (".conv3d." in k and k.count(".") == 2)), k     # the 1x1 convs of bilinear mode
    assert any(".bns.1.weight" in k for k in g1) and "out_conv.weight" in g1 and "block0.conv.conv3d_1.weight" in g1
    moved = [k for k, v in net.state_dict().items() if "running_mean" in k and not torch.equal(v, rm_before[k])]
    assert moved and all(".bns.1." in k for k in moved)
    # BatchNorm uses batch statistics: the second pass sees the same activations, so gradients scale exactly with the loss
    loss2, g2 = grads(2.0)
    assert abs(loss2 - loss1) <= 1e-5 * abs(loss1)
    assert g1.keys() == g2.keys()
    for k in g1:
        scale = float(g1[k].abs().max())
        if scale == 0.0:
            assert float(g2[k].abs().max()) == 0.0, k
            continue
        # bf16 gradients of the scaled loss round differently: relative to the tensor's scale
        assert float((g2[k] - 2 * g1[k]).abs().max()) <= 4e-2 * 2 * scale, k
        assert rel_l2(g2[k], 2 * g1[k]) < 2e-2, k


def test_full_size_filter_kernels_bit_exact():
    """configs[1] volume size (48x256x256): argmax pseudo labels, agreement weights (+ image-weight folding) and the
    disagreement count bit-exact against the numpy oracle (the MC-dropout statistics at this size:
    test_full_size_mc_uncertainty_and_five_class_filter below)."""
    from oracle import fpl_filter
    from fplplus_b200 import fpl
    r = np.random.Generator(np.random.PCG64(17))
    shape = (1, 2, 48, 256, 256)
    za = (r.standard_normal(shape) * 2).astype(np.float32)
    zb = (za + r.standard_normal(shape) * 0.7).astype(np.float32)
    za[0, :, 5, 7, :64] = 0.25                                   # ties: first index wins
    ta, tb = torch.from_numpy(za).to(DEV), torch.from_numpy(zb).to(DEV)
    la, lb, w, cnt = fpl.agreement_weight(ta, tb, image_weight=0.42)
    ra, rb = fpl_filter.pseudo_label(za)[0], fpl_filter.pseudo_label(zb)[0]
    ga, gb = la.cpu().numpy().reshape(ra.shape), lb.cpu().numpy().reshape(rb.shape)
    # labels are the argmax of fp32 softmax probabilities on both sides; the only freedom left is the last-ulp rounding of
    # expf (CUDA vs numpy) for logits inside the ~1e-7 band where probabilities collapse to a tie
    for got, ref, z in ((ga, ra, za), (gb, rb, zb)):
        bad = got != ref
        assert bad.sum() <= 2
        assert np.all(np.abs(z[0, 0] - z[0, 1])[bad] <= 2e-7)
    np.testing.assert_array_equal(ga, ra)                       # za has no such voxel besides the planted exact ties
    ref_w = fpl_filter.agreement_weight(ga, gb)
    np.testing.assert_array_equal(w.cpu().numpy().reshape(ref_w.shape),
                                  fpl_filter.set_weight_(np.float32(0.42), ref_w.astype(np.float32)))
    assert int(cnt) == int((ga != gb).sum())
    assert np.array_equal(ra[5, 7, :64], np.zeros(64, dtype=ra.dtype))


@pytest.mark.parametrize("c,k,shape", [(2, 6, (48, 256, 256)), (5, 6, (24, 160, 160)), (2, 12, (16, 64, 64)), (5, 16, (8, 32, 32))],
                         ids=["configs1_K6_C2_48x256x256", "mmwhs_K6_C5", "K12_streaming", "K16_C5_streaming"])
def test_full_size_mc_uncertainty_and_five_class_filter(c, k, shape):
    """K MC-dropout logits volumes -> (sum of variances, boundary count, uncertainty map) against the numpy restatement
    of agent_seg.py:911-929, at the configs[1] volume size with the reference's K = 6, for 5 classes (configs[4]) and
    for K > 8 (the streaming kernel; ADVICE r1: 9..16 passes used to fail after all forwards had run).  With 5 classes
    also the argmax labels / multi-class agreement weights (w = 1 - 0.5*[a != b])."""
    from oracle import fpl_filter
    from fplplus_b200 import fpl
    r = np.random.Generator(np.random.PCG64(100 + 10 * c + k))
    base = (r.standard_normal((1, c) + shape) * 2).astype(np.float32)
    passes = [(base + 0.6 * r.standard_normal(base.shape)).astype(np.float32) for _ in range(k)]
    assert k <= fpl.max_mc_passes()
    stats, umap = fpl.mc_uncertainty([torch.from_numpy(p).to(DEV) for p in passes], want_map=True)
    ref = fpl_filter.mc_uncertainty(passes)
    v, b = stats.tolist()
    ref_map = ref["uncertainty_map"].reshape(shape)
    near = int((np.abs(ref_map - 0.01) < 1e-6).sum())
    assert abs(int(b) - ref["boundary"]) <= near
    np.testing.assert_allclose(v, float(ref["vars"]), rtol=2e-5)
    np.testing.assert_allclose(umap.cpu().numpy(), ref_map, rtol=1e-4, atol=1e-7)
    got = fpl.finish_uncertainty(stats)
    assert abs(float(got) - float(ref["uncer_one"])) <= 3e-5 * abs(float(ref["uncer_one"])) + 1e-12
    with pytest.raises(ValueError):
        fpl.mc_uncertainty([torch.from_numpy(passes[0]).to(DEV)] * (fpl.max_mc_passes() + 1))
    if c > 2:
        za, zb = passes[0], passes[1]
        la, lb, w, cnt = fpl.agreement_weight(torch.from_numpy(za).to(DEV), torch.from_numpy(zb).to(DEV))
        ra, rb = fpl_filter.pseudo_label(za)[0], fpl_filter.pseudo_label(zb)[0]
        ga, gb = la.cpu().numpy().reshape(ra.shape), lb.cpu().numpy().reshape(rb.shape)
        assert (ga != ra).sum() <= 2 and (gb != rb).sum() <= 2        # last-ulp expf ties only (see the test above)
        assert ga.max() == c - 1
        np.testing.assert_array_equal(w.cpu().numpy().reshape(ra.shape).astype(np.float64),
                                      fpl_filter.agreement_weight_multiclass(ga, gb))
        assert int(cnt) == int((ga != gb).sum())
        np.testing.assert_array_equal(fpl.pseudo_label(torch.from_numpy(za).to(DEV)).cpu().numpy()[0], ga)


def test_full_size_window_stitching_identity():
    """Inferer over a 48x256x256 volume with a model that returns a fixed function of the window (its own voxels):
    the stitched, TTA-averaged result must equal that function of the volume, whatever the window overlap."""
    from fplplus_b200.inferer import Inferer

    class Toy(torch.nn.Module):
        def forward(self, x, domain_label=None):
            return torch.cat([x * 2.0 + 1.0, -x], 1)

    vol = torch.from_numpy(synth.synth_image(1, 1, (48, 256, 256), seed=41)).to(DEV)
    for stride in ([32, 128, 128], [16, 96, 96]):
        inf = Inferer({"class_num": 2, "sliding_window_enable": True, "sliding_window_size": [32, 128, 128],
                       "sliding_window_stride": stride, "tta_mode": 1})
        out = inf.run(Toy(), vol, torch.zeros(1, dtype=torch.long))
        want = torch.cat([vol * 2.0 + 1.0, -vol], 1)
        assert float((out - want).abs().max()) <= 1e-5


def test_shipped_vs_config_eval_logits_and_train_step():
    """SURVEY 8 f-1 at the authors' REAL configuration (config_dual/data_vs/vs_t1s_g.cfg:61-64,115-117): feature_chns
    32-64-128-256-512, conv_dims [2,2,3,3,3] (2.5-D: (1,3,3) convs, (1,2,2) pools / transposed convs on the first two
    levels), patches / windows 28x128x128.  Eval logits and one weighted Dice+CE train step against the oracle."""
    from oracle import losses, unet_dsbn
    from fplplus_b200.loss import CombinedLoss
    from fplplus_b200.registry import loss_dict
    params = dict(NET_PARAMS, feature_chns=[32, 64, 128, 256, 512], conv_dims=[2, 2, 3, 3, 3], dropout=[0.0] * 5)
    shape = (28, 128, 128)
    net, sd = _net(params)
    x = torch.from_numpy(synth.synth_image(1, 1, shape, seed=21))
    # eval, both domains
    net.eval()
    st = unet_dsbn.to_torch_state(sd, requires_grad=False)
    for dmn in (0, 1):
        with torch.no_grad():
            z = net(x.to(DEV), domain_label=dmn * torch.ones(1, dtype=torch.long)).cpu()
            ref = unet_dsbn.forward(st, x, dmn, params)
        e = rel_l2(z, ref)
        print("shipped VS config, domain %d: eval logits rel_l2 %.2e" % (dmn, e))
        assert e < 1e-2
    # one train step (batch 2, target domain, pixel-weighted 0.5 Dice + 0.5 CE)
    net.train()
    xb = torch.from_numpy(synth.synth_image(2, 1, shape, seed=22))
    lab = synth.synth_label(2, 2, shape, seed=22)
    y = torch.from_numpy(synth.one_hot(lab, 2))
    pw = torch.from_numpy(synth.synth_pixel_weight(lab, seed=22)[0])
    crit = CombinedLoss({"loss_type": ["DiceLoss", "CrossEntropyLoss"], "loss_weight": [0.5, 0.5]}, loss_dict)
    out = net(xb.to(DEV), domain_label=torch.ones(2, dtype=torch.long))
    loss = crit({"prediction": out, "ground_truth": y.to(DEV), "pixel_weight": pw.to(DEV)})
    loss.backward()
    st = unet_dsbn.to_torch_state(sd, requires_grad=True)
    ref_out = unet_dsbn.forward(st, xb, 1, params, bn_training=True)
    ref_loss = losses.combined_loss(ref_out, y, pw, 0.5, 0.5)
    ref_loss.backward()
    print("shipped VS config train step: logits rel_l2 %.2e loss %.6f oracle %.6f" % (
        rel_l2(out.detach().cpu(), ref_out.detach()), float(loss), float(ref_loss)))
    assert rel_l2(out.detach().cpu(), ref_out.detach()) < 1.5e-2
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss))
    named = dict(net.named_parameters())
    got_keys = {k for k, p in named.items() if p.grad is not None}
    ref_keys = {k for k, v in st.items() if v.requires_grad and v.grad is not None}
    assert got_keys == ref_keys                                  # 2-D members on levels 0-1, 3-D members below
    assert "block0.conv.conv2d_1.weight" in got_keys and "block2.conv.conv3d_1.weight" in got_keys
    for key in ("out_conv.weight", "out_conv.bias", "up4.conv.conv2d_2.weight", "up4.trans2d.weight"):
        e = rel_l2(named[key].grad.cpu(), st[key].grad)
        print("  d%-28s rel_l2 %.2e" % (key, e))
        assert e < 8e-2, (key, e)
    # running statistics of the selected domain
    sdn = net.state_dict()
    for key in ("block0.conv.bn2d1.bns.1.running_mean", "up4.conv.bn2d2.bns.1.running_var"):
        assert rel_l2(sdn[key].cpu(), st[key].detach()) < 2e-2, key
