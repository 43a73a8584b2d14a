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
