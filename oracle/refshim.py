"""Loader shim that executes the UNMODIFIED reference (sunshinnnn/DSMnet, Python 2.7 / PyTorch 0.3
era) under Python 3 / torch 2.x on CPU — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used only (a) by ``tests/golden/make_golden.py`` to freeze golden vectors from the reference's
own code and (b) by ``tests/test_oracle_vs_reference.py`` to pin ``oracle/ops.py`` live.  The
reference tree (``/root/reference`` or ``$DSMNET_REFERENCE``) exists only in the build container;
everything here degrades to "unavailable" elsewhere (``available()``).

Nothing is copied from the reference: its files are imported / exec'd where they lie.  The
shims are exactly the ones SURVEY.md App. A/B lists:
  * models/util_conv.py ends in a Python-2-only ``test()`` (``print net`` at :277) -> exec the
    source up to ``def test(`` only;
  * models import siblings by bare name (``from util_conv import ...``) -> sys.path entries;
  * models/iresnet.py:8 imports the wrong package name ``util.imwrap`` -> alias to ``utils``;
  * F.upsample / F.grid_sample changed their align_corners default after PyTorch 0.3 -> the
    shim pins align_corners=True (``pinned_torch`` context manager);
  * PSMNet.forward uses Py2 integer ``/`` and a hard-coded ``.cuda()``
    (stackhourglass.py:124-126) -> ``psmnet_forward_from_features`` calls the reference's own
    sub-modules in the order of :123-166 with ``//`` and no ``.cuda()``;
  * deconv3d_bn appends BatchNorm2d to a 3-D deconv (util_conv.py:176) -> swapped for
    BatchNorm3d of the same width in ``make_gcnet``.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("DSMNET_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "util_conv.py"))


_loaded: Dict[str, types.ModuleType] = {}


def _load() -> Dict[str, types.ModuleType]:
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)
    for p in (REF, REF + "/models", REF + "/models/psmnet", REF + "/losses"):
        if p not in sys.path:
            sys.path.insert(0, p)
    src = open(REF + "/models/util_conv.py", encoding="utf-8").read().split("\n")
    cut = next(i for i, line in enumerate(src) if line.startswith("def test("))
    util_conv = types.ModuleType("util_conv")
    util_conv.__file__ = REF + "/models/util_conv.py"
    exec(compile("\n".join(src[:cut]), util_conv.__file__, "exec"), util_conv.__dict__)
    sys.modules["util_conv"] = util_conv
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mp = types.ModuleType("matplotlib"); mp.use = lambda *a, **k: None
            sys.modules["matplotlib"] = mp
            sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    import importlib
    utils_pkg = importlib.import_module("utils")
    imwrap_mod = importlib.import_module("utils.imwrap")
    sys.modules.setdefault("util", utils_pkg)
    sys.modules.setdefault("util.imwrap", imwrap_mod)
    _loaded["util_conv"] = util_conv
    _loaded["imwrap"] = imwrap_mod
    _loaded["submodule"] = importlib.import_module("submodule")
    _loaded["stackhourglass"] = importlib.import_module("stackhourglass")
    _loaded["gcnet"] = importlib.import_module("gcnet")
    _loaded["dispnetcorr"] = importlib.import_module("dispnetcorr")
    _loaded["iresnet"] = importlib.import_module("iresnet")          # imports `util.imwrap` (sic, iresnet.py:8): aliased above
    return _loaded


@contextlib.contextmanager
def pinned_torch():
    """PyTorch<=0.3 semantics for the calls the reference makes: align_corners=True."""
    real_gs, real_up, real_interp = F.grid_sample, getattr(F, "upsample", None), F.interpolate

    def gs(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        return real_gs(inp, grid, mode=mode, padding_mode=padding_mode, align_corners=True)

    def interp(inp, size=None, scale_factor=None, mode="nearest", align_corners=None, **kw):
        if mode in ("linear", "bilinear", "trilinear", "bicubic"):
            align_corners = True
        return real_interp(inp, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners, **kw)

    F.grid_sample, F.interpolate, F.upsample = gs, interp, interp
    try:
        yield
    finally:
        F.grid_sample, F.interpolate = real_gs, real_interp
        if real_up is not None:
            F.upsample = real_up


# ---- op-level entry points (the reference's own code) ----------------------------------------

def corr1d(fL, fR, D, stride=1, kernel_size=1):
    m = _load()["util_conv"].Corr1d(kernel_size=kernel_size, stride=stride, D=D)
    return m(fL, fR)


def imwrap(im_src, disp, fliplr=False, LeftTop=(0, 0), scale_factor=1, seed=0):
    """Runs imwrap_BCHW as is; returns (out, delt) where delt is the value it drew (imwrap.py:70)."""
    mod = _load()["imwrap"]
    torch.manual_seed(seed)
    delt = float(1e-4 * (torch.rand(1)[0] + 0.1))
    torch.manual_seed(seed)
    with pinned_torch():
        out = mod.imwrap_BCHW(im_src, disp, fliplr, list(LeftTop), scale_factor)
    return out, delt


def psm_volume(fL, fR, maxdisp):
    """The loop of stackhourglass.py:124-133 re-typed with '//' and without .cuda() (A3/A4)."""
    C = fL.size(1)
    D = maxdisp // 4
    cost = torch.zeros(fL.size(0), C * 2, D, fL.size(2), fL.size(3))
    for i in range(D):
        if i > 0:
            cost[:, :C, i, :, i:] = fL[:, :, :, i:]
            cost[:, C:, i, :, i:] = fR[:, :, :, :-i]
        else:
            cost[:, :C, i, :, :] = fL
            cost[:, C:, i, :, :] = fR
    return cost.contiguous()


def make_psmnet(maxdisp=192, seed=0):
    mods = _load()
    torch.manual_seed(seed)
    net = mods["stackhourglass"].PSMNet(maxdisp)
    net.eval()
    return net


def psmnet_3d(net, cost):
    """dres0..classif3 exactly as wired at stackhourglass.py:135-149, on the reference's modules."""
    sh = _load()["stackhourglass"]
    cost0 = net.dres0(cost)
    cost0 = sh.myadd_3d(net.dres1(cost0), cost0)
    out1, pre1, post1 = net.dres2(cost0, None, None)
    out1 = sh.myadd_3d(out1, cost0)
    out2, pre2, post2 = net.dres3(out1, pre1, post1)
    out2 = sh.myadd_3d(out2, cost0)
    out3, pre3, post3 = net.dres4(out2, pre1, post2)
    out3 = sh.myadd_3d(out3, cost0)
    cost1 = net.classif1(out1)
    cost2 = net.classif2(out2) + cost1
    cost3 = net.classif3(out3) + cost2
    return cost1, cost2, cost3


def psmnet_heads(net, costs, maxdisp, H, W):
    """stackhourglass.py:152-166 with the reference's disparityregression module."""
    sub = _load()["submodule"]
    preds = []
    with pinned_torch():
        for c in costs:
            c = F.upsample(c, [maxdisp, H, W], mode="trilinear")
            c = torch.squeeze(c, 1)
            p = F.softmax(c, dim=1)
            preds.append(sub.disparityregression(maxdisp)(p))
    return preds


def psmnet_forward_from_features(net, fL, fR, maxdisp, H, W):
    cost = psm_volume(fL, fR, maxdisp)
    c1, c2, c3 = psmnet_3d(net, cost)
    p1, p2, p3 = psmnet_heads(net, (c1, c2, c3), maxdisp, H, W)
    return [p3, p2, p1]


def make_gcnet(maxdisparity=192, seed=0):
    mods = _load()
    torch.manual_seed(seed)
    net = mods["gcnet"].gcnet(maxdisparity)
    net.D = int(net.D)
    for name in ("l33", "l34", "l35", "l36"):
        seq = getattr(net.layer3d, name)
        for i, m in enumerate(seq):
            if isinstance(m, nn.BatchNorm2d):
                seq[i] = nn.BatchNorm3d(m.num_features)
    return net


def gc_volume(fL, fR, D, right=False):
    """gcnet.py:131-135 (left) / :157-164 (right-reference) loops as written there."""
    n, Fc, h, w = fL.shape
    if not right:
        xL = torch.zeros(n, Fc * 2, D, h, w)
        xL[:, :, 0] = torch.cat([fL, fR], 1)
        for i in range(1, D):
            xL[:, :Fc, i] = fL
            xL[:, Fc:, i, :, i:] = fR[:, :, :, :-i]
        return xL
    xR = torch.zeros(n, Fc * 2, D, h, w)
    xR[:, :, 0] = torch.cat([fR, fL], 1)
    for i in range(1, D):
        xR[:, :Fc, i] = fR
        xR[:, Fc:, i, :, :-i] = fL[:, :, :, i:]
    return xR


def state_dict_3d(net) -> Dict[str, torch.Tensor]:
    """Parameters/buffers of the 3-D part of a reference PSMNet, reference names."""
    return {k: v.detach().clone() for k, v in net.state_dict().items()
            if k.split(".")[0] in ("dres0", "dres1", "dres2", "dres3", "dres4", "classif1", "classif2", "classif3")
            and not k.endswith("num_batches_tracked")}


# ------------------------------------------------------------------------------------------------
# losses/loss.py (self-supervised training config): Python-2 prints and PyTorch-0.3 idioms are
# patched textually before exec — nothing else is changed.
# ------------------------------------------------------------------------------------------------

def load_losses():
    """The reference's `losses.loss` module (class `losses`), executable under Python 3 / torch 2.x."""
    import re
    mods = _load()
    if "ref_loss" in mods:
        return mods["ref_loss"]
    if "matplotlib" not in sys.modules:                       # utils/utils.py imports it at module level
        mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt; mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    import SSIM as ref_ssim                                   # losses/SSIM.py, unmodified (REF/losses is on sys.path)
    src = open(REF + "/losses/loss.py", encoding="utf-8").read()
    src = re.sub(r"^(\s*)print [^\n]*$", r"\1pass", src, flags=re.M)                       # 8 Py2 print statements
    src = src.replace(".data[0]", ".item()")                                                 # 0-dim tensor indexing
    src = re.sub(r"\(\((.+?)\) \+ mask_ap\)\.detach\(\) > 1", r"((\1) & mask_ap).detach()", src)   # byte '+' used as AND
    src = src.replace("mask2 = (disp_delt<3) - mask1", "mask2 = (disp_delt<3) & ~mask1")     # byte '-' used as AND-NOT
    src = src.replace("from utils.utils import imsplot_tensor", "imsplot_tensor = None")
    m = types.ModuleType("ref_loss")
    m.__file__ = REF + "/losses/loss.py"
    exec(compile(src, m.__file__, "exec"), m.__dict__)
    mods["ref_loss"] = m
    return m


def selfsup_loss(loss_name, scale_disps, dispLs, dispL1s, imL, imR_src, imL1, imR1_src, LeftTop, weight_levels, seed=0):
    """losses(loss_name)(...) exactly as stereo_selfsupervised.py:88-95 calls it; the global RNG is seeded so that the
    `delt` draws of the imwrap calls (imwrap.py:70) are reproducible.  Returns the scalar loss tensor."""
    m = load_losses()
    L = m.losses(loss_name=loss_name, count_levels=len(weight_levels))
    L.weight_levels = list(weight_levels)
    torch.manual_seed(seed)
    with pinned_torch():
        return L.lossesfun(imR_src, imL, dispLs, scale_disps, list(LeftTop), imR1_src, imL1, dispL1s, scale_disps, list(LeftTop))


# ------------------------------------------------------------------------------------------------
# TRAIN-mode runs of the reference's own 3-D stacks (batch-statistics BatchNorm, autograd): what pins
# oracle.ops.psmnet_hotpath_train / gcnet_hotpath_train (tests/golden/make_golden_train.py, tests/test_oracle_vs_reference.py)
# ------------------------------------------------------------------------------------------------

def psmnet_train_from_features(net, fL, fR, maxdisp, H, W):
    """stackhourglass.py:123-168 from the feature maps on with the reference's modules under net.train(); fL / fR and the
    parameters may require grad (the volume loop of :124-133 is differentiable through its slice assignments)."""
    net.train()
    return psmnet_forward_from_features(net, fL, fR, maxdisp, H, W)


def gcnet_train_from_features(net, fL, fR):
    """gcnet.forward (:126-137) from the feature maps on with the reference's layer3d under net.train()."""
    net.train()
    vol = gc_volume(fL, fR, int(net.D))
    with pinned_torch():
        return net.layer3d(vol, "train")
