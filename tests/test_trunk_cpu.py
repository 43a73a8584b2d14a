"""CPU: whole-model drop-ins against the reference's state-dict layout, and the 2-D trunk restatements (oracle + the
modules' stock-PyTorch graphs) against outputs of the reference's own modules (tests/golden/make_golden_models.py)."""
import json
import os

import pytest
import torch

import oracle.ops as O
from conftest import GOLDEN, load_golden, rel_err
from helpers import trunk_params_from_golden


def _keys(name):
    return json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))[name]


@pytest.mark.parametrize("name", ["psmnet", "gcnet", "dispnetcorr", "iresnet"])
def test_state_dict_layout_matches_reference(name):
    """a reference checkpoint (`weight_best.pkl`, stereo.py:81-83) loads with strict=True: same keys, same shapes"""
    if name == "psmnet":
        from dsmnet_b200.psmnet import PSMNet as M
    elif name == "gcnet":
        from dsmnet_b200.gcnet import gcnet as M
    elif name == "dispnetcorr":
        from dsmnet_b200.dispnetcorr import dispnetcorr as M
    else:
        from dsmnet_b200.iresnet import iresnet as M
    mine = {k: list(v.shape) for k, v in M(192).state_dict().items()}
    ref = _keys(name)
    assert set(mine) == set(ref), (sorted(set(mine) - set(ref))[:5], sorted(set(ref) - set(mine))[:5])
    assert all(mine[k] == ref[k] for k in ref)


def test_gcnet_feature3d_keys_are_the_layer3d_part():
    from dsmnet_b200.gcnet import feature3d
    ref = {k[len("layer3d."):]: v for k, v in _keys("gcnet").items() if k.startswith("layer3d.")}
    mine = {k: list(v.shape) for k, v in feature3d(32).state_dict().items()}
    assert mine == ref and len(ref) > 100


def test_psmnet_trunk_oracle_and_module_vs_reference_golden():
    from dsmnet_b200.psmnet import feature_extraction
    g = load_golden("psmnet_trunk")
    m = feature_extraction().eval()
    p = trunk_params_from_golden(g, m)
    taps = {}
    with torch.no_grad():
        out = O.psmnet_feature_extraction(p, g["x"], taps=taps)
    assert out.shape == g["out"].shape and rel_err(out, g["out"]) < 1e-4
    for mine, ref in (("firstconv", "stage_firstconv"), ("layer1", "stage_layer1"), ("raw", "stage_layer2"),
                      ("layer3", "stage_layer3"), ("skip", "stage_layer4")):
        assert rel_err(taps[mine][:, :, ::4, ::4], g[ref]) < 1e-4
    missing = m.load_state_dict(p, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    with torch.no_grad():
        assert rel_err(m(g["x"]), g["out"]) < 1e-4           # the module's stock graph (CPU / autograd route)


def test_gcnet_trunk_oracle_and_module_vs_reference_golden():
    from dsmnet_b200.gcnet import feature2d
    g = load_golden("gcnet_trunk")
    m = feature2d(32).eval()
    p = trunk_params_from_golden(g, m)
    with torch.no_grad():
        out = O.gcnet_feature2d(p, g["x"])
    assert out.shape == g["out"].shape and rel_err(out, g["out"]) < 1e-4
    missing = m.load_state_dict(p, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    with torch.no_grad():
        assert rel_err(m(g["x"]), g["out"]) < 1e-4
