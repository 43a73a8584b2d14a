"""Freeze TRAIN-mode golden vectors (loss, predictions, gradients, running statistics) from the reference's OWN
modules under .train() — run in the build container only:

    python tests/golden/make_golden_train.py        # needs /root/reference (or $DSMNET_REFERENCE)

These pin oracle.ops.psmnet_hotpath_train / gcnet_hotpath_train (tests/test_oracle_golden.py) — the yardstick of the
GPU training-parity tests — to the reference: models/psmnet/stackhourglass.py:123-168 and models/gcnet.py:65-111,126-137
with batch-statistics BatchNorm3d, through torch autograd.  The odd-sized cases exercise the crop-to-min skip adds
(myadd_3d, stackhourglass.py:10-20; myAdd3d, util_fun.py:41-51) with BatchNorm statistics over the uncropped tensor.
Kept separate from make_golden.py so that the existing fixtures (whose values depend on that script's RNG order) stay
byte-identical."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle.ops as O          # noqa: E402
import oracle.refshim as R      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PSM_GRADS = ("dres0.0.0.weight", "dres1.2.0.weight", "dres2.conv1.0.0.weight", "dres2.conv2.0.weight", "dres3.conv5.0.weight",
             "dres4.conv6.0.weight", "classif1.0.0.weight", "classif3.2.weight", "dres2.conv2.1.weight", "dres4.conv6.1.bias")
GC_GRADS = ("l19.0.weight", "l21.0.weight", "l22.0.weight", "l30.0.weight", "l32.0.weight", "l33.0.weight", "l36.0.weight",
            "l37.weight", "l35.1.weight", "l34.1.bias")


def save(name, **arrays):
    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote %-28s %6.1f KB" % (name + ".npz", os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


def grads(named, keys):
    """gradients of the listed parameters; big ones are stored as every 5th element of the flattened tensor (suffix __s5)"""
    out = {}
    for k in keys:
        g = named[k].grad
        name = "g_" + k.replace(".", "_")
        if g.numel() > 60000:
            out[name + "__s5"] = g.flatten()[::5].clone()
        else:
            out[name] = g
    return out


def psm_case(name, B, h, w, maxdisp, seed):
    rs = np.random.RandomState(seed)
    fL = torch.from_numpy(rs.standard_normal((B, 32, h, w)).astype(np.float32))
    fR = torch.from_numpy(rs.standard_normal((B, 32, h, w)).astype(np.float32))
    H, W = 4 * h, 4 * w
    gt = torch.from_numpy(rs.uniform(0, maxdisp * 0.5, size=(B, H, W)).astype(np.float32))
    params = O.psmnet_random_params(seed=seed + 1)
    net = R.make_psmnet(maxdisp, seed=0)
    missing = net.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys
    a = fL.clone().requires_grad_(); b = fR.clone().requires_grad_()
    preds = R.psmnet_train_from_features(net, a, b, maxdisp, H, W)                  # [pred3, pred2, pred1]
    loss = sum(wt * (p - gt).abs().mean() for wt, p in zip((1.0, 0.7, 0.5), preds))   # the reference's weighted L1 pyramid
    loss.backward()
    named = dict(net.named_parameters())
    sd = net.state_dict()
    save(name, fL=fL, fR=fR, gt=gt, maxdisp=maxdisp, seed=seed + 1, loss=loss.detach(),
         pred3=preds[0], pred2=preds[1], pred1=preds[2], gL=a.grad, gR=b.grad,
         rm_dres0_0=sd["dres0.0.1.running_mean"], rv_dres0_0=sd["dres0.0.1.running_var"],
         rm_conv5=sd["dres3.conv5.1.running_mean"], rv_conv5=sd["dres3.conv5.1.running_var"],
         **grads(named, PSM_GRADS))


def gc_case(name, B, h, w, maxdisp, seed):
    rs = np.random.RandomState(seed)
    fL = torch.from_numpy(rs.standard_normal((B, 32, h, w)).astype(np.float32))
    fR = torch.from_numpy(rs.standard_normal((B, 32, h, w)).astype(np.float32))
    gt = torch.from_numpy(rs.uniform(0, maxdisp * 0.5, size=(B, 1, 2 * h, 2 * w)).astype(np.float32))
    params = O.gcnet_random_params(seed=seed + 1)
    net = R.make_gcnet(maxdisp, seed=0)
    missing = net.load_state_dict({"layer3d." + k: v for k, v in params.items()}, strict=False)
    assert not missing.unexpected_keys
    a = fL.clone().requires_grad_(); b = fR.clone().requires_grad_()
    disp = R.gcnet_train_from_features(net, a, b)
    loss = (disp - gt[:, :, :disp.shape[2], :disp.shape[3]]).abs().mean()
    loss.backward()
    named = dict(net.layer3d.named_parameters())
    sd = net.layer3d.state_dict()
    save(name, fL=fL, fR=fR, gt=gt, maxdisp=maxdisp, seed=seed + 1, loss=loss.detach(), disp=disp, gL=a.grad, gR=b.grad,
         rm_l33=sd["l33.1.running_mean"], rv_l33=sd["l33.1.running_var"],
         **grads(named, GC_GRADS))


def main():
    assert R.available(), "reference tree not found"
    psm_case("psmnet_train", 2, 8, 16, 32, 300)          # even sizes: D 8 -> 4 -> 2
    psm_case("psmnet_train_odd", 1, 7, 11, 24, 310)      # D 6 -> 3 -> 2: conv5 / conv6 outputs are cropped at the adds
    gc_case("gcnet_train", 2, 16, 32, 32, 320)           # D 16 -> 8 -> 4 -> 2 -> 1
    gc_case("gcnet_train_odd", 1, 18, 22, 36, 330)       # D 18 -> 9 -> 5 -> 3 -> 2: every skip add crops


if __name__ == "__main__":
    main()
