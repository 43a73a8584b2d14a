"""CPU: the graph (host logic) of the whole-model drop-ins against pyramids produced by the reference's own modules, with the
oracle's CPU ops injected at the hot-path call sites (the CUDA ops have no CPU form)."""
import numpy as np
import torch

import oracle.ops as O
from conftest import load_golden
from helpers import random_state_dict


def test_iresnet_graph_matches_reference():
    """BASELINE config 4: iresnet.forward (models/iresnet.py:87-200): two Corr1d call sites and the imwrap feature warp"""
    from dsmnet_b200.iresnet import iresnet
    g = load_golden("iresnet_forward")

    def warp_cpu(src, disp):
        delt = float(1e-4 * (torch.rand(1)[0] + 0.1))
        return O.imwrap(src, disp, False, (0, 0), 1, delt)

    m = iresnet(192, True, lambda a, b, D, s, k: O.corr1d(a, b, D, s, k), warp_cpu).eval()
    m.load_state_dict(random_state_dict(m, g["seed"]), strict=True)
    with torch.no_grad():
        torch.manual_seed(g["rng_seed"])
        scales, outs = m(g["imL"], g["imR"], "test")
    assert scales == [int(x) for x in g["scales"]]
    for i, o in enumerate(outs):
        assert torch.allclose(o, g["out%d" % i], rtol=1e-5, atol=1e-6)
