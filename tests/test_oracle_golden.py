"""CPU: the oracle (oracle/ops.py) against the golden vectors frozen from the reference's own code."""
import hashlib

import numpy as np
import pytest
import torch

import oracle.ops as O
from conftest import load_golden, rel_err

CORR = ["corr1d_dispnetc", "corr1d_iresnet2", "corr1d_d_gt_w"]


@pytest.mark.parametrize("name", CORR)
def test_corr1d(name):
    g = load_golden(name)
    out = O.corr1d(g["fL"], g["fR"], g["D"], g["stride"], g["kernel_size"])
    assert torch.equal(out, g["out"])          # same op sequence as the reference -> bit-exact
    if g["kernel_size"] == 1:
        gL, gR = O.corr1d_grads(g["gout"], g["fL"], g["fR"], g["stride"])
        assert rel_err(gL, g["gL"]) < 1e-5 and rel_err(gR, g["gR"]) < 1e-5


def test_volumes():
    g = load_golden("volume_psm")
    assert torch.equal(O.concat_volume(g["fL"], g["fR"], g["maxdisp"] // 4, "psm"), g["out"])
    g = load_golden("volume_gc")
    assert torch.equal(O.concat_volume(g["fL"], g["fR"], g["D"], "gc"), g["out"])
    g = load_golden("volume_gc_right")
    assert torch.equal(O.concat_volume(g["fL"], g["fR"], g["D"], "gc_right"), g["out"])
    g = load_golden("volume_gc_bwd")
    gL, gR = O.concat_volume_grads(g["gout"], g["D"], "gc")
    assert rel_err(gL, g["gL"]) < 1e-6 and rel_err(gR, g["gR"]) < 1e-6


def test_heads():
    g = load_golden("head_psm")
    pred = O.upsample_softargmin(g["cost_lr"].squeeze(1), g["size"], True)
    assert rel_err(pred, g["pred"]) < 1e-6
    assert rel_err(O.disparity_regression(g["prob"]), g["pred"]) < 1e-6
    g = load_golden("head_gc")
    assert rel_err(O.softargmin(g["x37"].squeeze(1), -1.0).unsqueeze(1), g["pred"]) < 1e-6


IMWRAP = ["imwrap_plain", "imwrap_fliplr", "imwrap_lefttop", "imwrap_scale2", "imwrap_intdisp", "imwrap_oob"]


@pytest.mark.parametrize("name", IMWRAP)
def test_imwrap(name):
    g = load_golden(name)
    out = O.imwrap(g["src"], g["disp"], g["fliplr"], tuple(g["LeftTop"]), g["scale_factor"], g["delt"])
    assert torch.equal(out, g["out"])
    # independent numpy closed form: values to 1e-6, and it yields the integer sampling indices
    _, _, h0, w0 = g["src"].shape
    _, _, h, w = g["disp"].shape
    row, col = O.imwrap_rowcol(h0, w0, h, w, tuple(g["LeftTop"]), g["scale_factor"])
    cf, x0, y0 = O.imwrap_closed_form(g["src"].numpy(), g["disp"].numpy(), row.numpy(), col.numpy(), g["fliplr"], g["delt"])
    assert np.abs(cf - g["out"].numpy()).max() < 2e-6


def test_psmnet_hotpath():
    g = load_golden("psmnet_hotpath")
    cost = O.concat_volume(g["fL"], g["fR"], g["maxdisp"] // 4, "psm")
    params = O.psmnet_random_params(seed=g["seed"], calibrate_on=cost)
    h = hashlib.sha256()
    for k in sorted(params):
        if params[k].dim() == 5:
            h.update(k.encode()); h.update(params[k].numpy().tobytes())
    assert h.hexdigest() == g["params_sha256"], "synthetic weights are not the ones the fixture was made with"
    c1, c2, c3 = O.psmnet_aggregate(params, cost)
    for mine, ref in ((c1, g["cost1"]), (c2, g["cost2"]), (c3, g["cost3"])):
        assert rel_err(mine, ref) < 1e-4
    preds = O.psmnet_hotpath(params, g["fL"], g["fR"], g["maxdisp"], (g["H"], g["W"]))
    for mine, ref in zip(preds, (g["pred3"], g["pred2"], g["pred1"])):
        assert (mine - ref).abs().max() < 2e-3       # px; fp32 BN folded vs unfolded
        assert (mine - ref).abs().mean() < 1e-4


def test_dispnetc_forward():
    """BASELINE config 1: the reference's DispNetC (with its Corr1d layer) vs the oracle restatement, all 7 levels"""
    g = load_golden("dispnetc_forward")
    outs = O.dispnetc_forward(O.dispnetc_random_params(seed=g["seed"]), g["imL"], g["imR"], "test")
    for i, o in enumerate(outs):
        assert o.shape == g["out%d" % i].shape and rel_err(o, g["out%d" % i]) < 1e-5


def test_gcnet_hotpath():
    """the reference's feature3d (19-layer 3-D enc-dec + soft-argmin of -cost) vs the oracle restatement"""
    g = load_golden("gcnet_hotpath")
    cost = O.concat_volume(g["fL"], g["fR"], g["maxdisp"] // 2, "gc")
    params = O.gcnet_random_params(seed=g["seed"], calibrate_on=cost)
    h = hashlib.sha256()
    for k in sorted(params):
        if params[k].dim() == 5:
            h.update(k.encode()); h.update(params[k].numpy().tobytes())
    assert h.hexdigest() == g["params_sha256"], "synthetic weights are not the ones the fixture was made with"
    disp = O.gcnet_hotpath(params, g["fL"], g["fR"], g["maxdisp"])
    assert disp.shape == g["disp"].shape
    assert (disp - g["disp"]).abs().max() < 2e-3 and (disp - g["disp"]).abs().mean() < 1e-4


@pytest.mark.parametrize("name,transposed,stride", [("gc_conv_s2", False, 2), ("gc_deconv", True, 2)])
def test_gc_layers(name, transposed, stride):
    g = load_golden(name)
    bn = {"weight": g["bn_weight"], "bias": g["bn_bias"], "running_mean": g["bn_mean"], "running_var": g["bn_var"]}
    cout = g["weight"].shape[1] if transposed else g["weight"].shape[0]
    scale, shift = O.fold_bn(cout, bn, g["bias"])
    y = O.conv3d_block(g["x"], g["weight"], scale, shift, stride, transposed, None, True)
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) < 1e-5


# ---- TRAIN mode: the oracle's batch-statistics restatements against the reference's own modules under .train() ---------
# (fixtures: tests/golden/make_golden_train.py).  This is what pins the yardstick of the GPU training-parity tests.

@pytest.mark.parametrize("name", ["psmnet_train", "psmnet_train_odd"])
def test_psmnet_train_oracle_vs_reference_golden(name):
    from helpers import PSM_TRAIN_GRADS, golden_grad, cosine, psm_train_loss
    g = load_golden(name)
    params = O.psmnet_random_params(seed=g["seed"])
    req = lambda k, v: v.dim() == 5 or k.endswith(".1.weight") or k.endswith(".1.bias")
    pr = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    a = g["fL"].clone().requires_grad_(); b = g["fR"].clone().requires_grad_()
    H, W = g["gt"].shape[-2:]
    preds = O.psmnet_hotpath_train(pr, a, b, g["maxdisp"], (H, W))
    loss = psm_train_loss(preds, g["gt"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    for mine, ref in zip(preds, (g["pred3"], g["pred2"], g["pred1"])):
        assert mine.shape == ref.shape and (mine - ref).abs().max() < 5e-3
    for mine, ref in ((a.grad, g["gL"]), (b.grad, g["gR"])):
        assert cosine(mine, ref) > 0.9999 and rel_err(mine, ref) < 2e-2
    for k in PSM_TRAIN_GRADS:
        mine, ref = golden_grad(g, k, pr[k].grad)
        assert cosine(mine, ref) > 0.9999 and rel_err(mine, ref) < 2e-2, k


@pytest.mark.parametrize("name", ["gcnet_train", "gcnet_train_odd"])
def test_gcnet_train_oracle_vs_reference_golden(name):
    from helpers import GC_TRAIN_GRADS, golden_grad, cosine
    g = load_golden(name)
    params = O.gcnet_random_params(seed=g["seed"])
    req = lambda k, v: v.dim() == 5 or k.endswith(".bias") or k.endswith(".1.weight")
    pr = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    a = g["fL"].clone().requires_grad_(); b = g["fR"].clone().requires_grad_()
    disp = O.gcnet_hotpath_train(pr, a, b, g["maxdisp"])
    assert disp.shape == g["disp"].shape and (disp - g["disp"]).abs().max() < 5e-3
    gt = g["gt"][:, :, :disp.shape[2], :disp.shape[3]]
    loss = (disp - gt).abs().mean()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    for mine, ref in ((a.grad, g["gL"]), (b.grad, g["gR"])):
        assert cosine(mine, ref) > 0.9999 and rel_err(mine, ref) < 2e-2
    for k in GC_TRAIN_GRADS:
        mine, ref = golden_grad(g, k, pr[k].grad)
        assert cosine(mine, ref) > 0.9999 and rel_err(mine, ref) < 2e-2, k
