"""Whole-model fixtures from the reference's OWN modules (run in the build container only):

    python tests/golden/make_golden_models.py

  * state_dict_keys.json — parameter / buffer names and shapes of the reference's PSMNet, gcnet, dispnetcorr and iresnet:
    what `model.load_state_dict(torch.load("weight_best.pkl")["state_dict"])` (stereo.py:81-83) has to match.
  * psmnet_trunk.npz — feature_extraction (models/psmnet/submodule.py:65-140) on a 256x320 image: synthetic parameters
    (oracle.ops.trunk_random_params, numpy RNG), BatchNorm running statistics calibrated by one train-mode pass of the
    reference with momentum 1 (stored in the fixture), the eval-mode output and strided samples of the intermediate stages.
  * gcnet_trunk.npz — feature2d (models/gcnet.py:14-29) on a 96x160 image, same recipe.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle.ops as O          # noqa: E402
import oracle.refshim as R      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote %-28s %6.1f KB" % (name + ".npz", os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


def calibrated(module, params, x):
    """load `params`, set every BatchNorm2d's running statistics to the batch statistics of `x` (one train-mode pass with
    momentum 1, as a freshly trained net would have), return the full eval-mode state dict"""
    module.load_state_dict(params, strict=False)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    module.train()
    with torch.no_grad(), R.pinned_torch():
        module(x)
    module.eval()
    return {k: v.detach().clone() for k, v in module.state_dict().items() if not k.endswith("num_batches_tracked")}


def main():
    assert R.available(), "reference tree not found"
    mods = R._load()
    keys = {}
    for name, net in (("psmnet", R.make_psmnet(192)), ("gcnet", R.make_gcnet(192)),
                      ("dispnetcorr", mods["dispnetcorr"].dispnetcorr(192)), ("iresnet", mods["iresnet"].iresnet(192))):
        keys[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("wrote state_dict_keys.json", {k: len(v) for k, v in keys.items()})

    # ---- PSMNet trunk ------------------------------------------------------------------------------------------------
    fe = mods["submodule"].feature_extraction()
    shapes = {k: tuple(v.shape) for k, v in fe.state_dict().items()}
    rs = np.random.RandomState(77)
    x = torch.from_numpy(rs.standard_normal((1, 3, 256, 320)).astype(np.float32))
    sd = calibrated(fe, O.trunk_random_params(shapes, seed=41), x)
    stages = {}
    hooks = [getattr(fe, n).register_forward_hook(lambda m, i, o, n=n: stages.__setitem__(n, o.detach().clone()))
             for n in ("firstconv", "layer1", "layer2", "layer3", "layer4")]
    with torch.no_grad(), R.pinned_torch():
        y = fe(x)
    for h in hooks:
        h.remove()
    bn = {k.replace(".", "__"): v for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    save("psmnet_trunk", x=x, seed=41, out=y, **{"stage_" + n: t[:, :, ::4, ::4].contiguous() for n, t in stages.items()}, **bn)

    # ---- GC-Net trunk ------------------------------------------------------------------------------------------------
    g2 = R.make_gcnet(192).layer2d
    shapes = {k: tuple(v.shape) for k, v in g2.state_dict().items()}
    x = torch.from_numpy(rs.standard_normal((1, 3, 96, 160)).astype(np.float32))
    sd = calibrated(g2, O.trunk_random_params(shapes, seed=43), x)
    with torch.no_grad():
        y = g2(x)
    bn = {k.replace(".", "__"): v for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    save("gcnet_trunk", x=x, seed=43, out=y, **bn)


if __name__ == "__main__":
    main()
