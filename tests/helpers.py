"""Test helpers shared by the CPU and GPU suites (not product code)."""
import math

import numpy as np
import torch


def random_state_dict(module: torch.nn.Module, seed: int, small=("pr", "r_res")):
    """A deterministic, host-independent state_dict for `module` (numpy MT19937, not torch's CPU RNG, whose stream depends
    on the SIMD width): conv / deconv weights ~ N(0, 2/(k*k*Cout)), scaled by 0.1 for the prediction heads, small biases."""
    rs = np.random.RandomState(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if v.dim() == 4:
            cout = v.shape[0]
            std = math.sqrt(2.0 / (v.shape[2] * v.shape[3] * cout))
            if k.split(".")[0].startswith(small):
                std *= 0.1
            sd[k] = torch.from_numpy(rs.standard_normal(size=tuple(v.shape)).astype(np.float32) * np.float32(std))
        else:
            sd[k] = torch.from_numpy(rs.standard_normal(size=tuple(v.shape)).astype(np.float32) * np.float32(0.01))
    return sd


# ---- train-mode golden fixtures (tests/golden/make_golden_train.py) -------------------------------------------------

PSM_TRAIN_GRADS = ("dres0.0.0.weight", "dres1.2.0.weight", "dres2.conv1.0.0.weight", "dres2.conv2.0.weight", "dres3.conv5.0.weight",
                   "dres4.conv6.0.weight", "classif1.0.0.weight", "classif3.2.weight", "dres2.conv2.1.weight", "dres4.conv6.1.bias")
GC_TRAIN_GRADS = ("l19.0.weight", "l21.0.weight", "l22.0.weight", "l30.0.weight", "l32.0.weight", "l33.0.weight", "l36.0.weight",
                  "l37.weight", "l35.1.weight", "l34.1.bias")


def golden_grad(g, key, mine):
    """(mine, golden) for parameter `key` of a train fixture; big gradients were stored as every 5th flattened element."""
    name = "g_" + key.replace(".", "_")
    if name in g:
        return mine.detach().cpu().flatten(), g[name].flatten()
    return mine.detach().cpu().flatten()[::5], g[name + "__s5"].flatten()


def cosine(u, v):
    return float(torch.nn.functional.cosine_similarity(u.flatten().double(), v.flatten().double(), dim=0))


def psm_train_loss(preds, gt):
    """the reference's weighted L1 pyramid over [pred3, pred2, pred1]"""
    return sum(wt * (p - gt).abs().mean() for wt, p in zip((1.0, 0.7, 0.5), preds))


# ---- 2-D trunk fixtures (tests/golden/make_golden_models.py) ---------------------------------------------------------

def trunk_params_from_golden(g, module):
    """the fixture's parameters: oracle.ops.trunk_random_params(seed) over `module`'s state_dict layout + the calibrated
    BatchNorm running statistics stored in the fixture"""
    import oracle.ops as O
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    p = O.trunk_random_params(shapes, seed=g["seed"])
    for k in list(p):
        if k.endswith("running_mean") or k.endswith("running_var"):
            p[k] = g[k.replace(".", "__")].clone()
    return p
