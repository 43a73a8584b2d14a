"""CPU: checkpoint and image I/O of the reference (dsmnet_b200/io.py)."""
import io as _io
import os

import numpy as np
import pytest
import torch

import oracle.refshim as R
from dsmnet_b200 import io


def test_pfm_roundtrip(tmp_path):
    rs = np.random.RandomState(0)
    for shape in ((7, 11), (5, 9, 3)):
        img = rs.standard_normal(shape).astype(np.float32)
        p = str(tmp_path / "x.pfm")
        io.save_pfm(p, img, scale=1)
        back, scale = io.load_pfm(p)
        assert scale == 1.0 and back.dtype == np.float32 and np.array_equal(back, img)
        assert np.array_equal(io.imread(p), img)


def test_png_decoder_matches_pil(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(1)
    smooth = (np.add.outer(np.arange(40), np.arange(61))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)   # filters 1-4 get used
    noise = rs.randint(0, 256, size=(23, 37, 3)).astype(np.uint8)
    grey16 = rs.randint(0, 65536, size=(19, 31)).astype(np.uint16)                                               # KITTI disparity maps
    for name, arr in (("s", smooth), ("n", noise), ("g", grey16)):
        p = str(tmp_path / (name + ".png"))
        Image.fromarray(arr).save(p)
        dec = io._png_decode(open(p, "rb").read())
        assert np.array_equal(dec, arr)
        assert np.array_equal(io.imread(p), arr)


def test_real_fixture_is_the_reference_pair():
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "kitti_pair.npz"))
    assert z["L"].shape == z["R"].shape == (372, 1240, 3) and z["L"].dtype == np.uint8
    ref = os.path.join(R.REF, "deploy", "10L.png")
    if os.path.isfile(ref):
        assert np.array_equal(io.imread(ref)[:372, :1240], z["L"])


@pytest.mark.parametrize("name", ["psmnet", "gcnet", "dispnetcorr", "iresnet"])
def test_reference_format_checkpoint_loads_strict(name, tmp_path):
    """a `weight_best.pkl` (stereo.py:80-83) written by the reference's own model class — when the reference tree is present,
    otherwise by the drop-in — loads with strict=True and restores every tensor"""
    torch.manual_seed(0)
    if R.available():
        mods = R._load()
        src = {"psmnet": lambda: R.make_psmnet(192), "gcnet": lambda: R.make_gcnet(192),
               "dispnetcorr": lambda: mods["dispnetcorr"].dispnetcorr(192), "iresnet": lambda: mods["iresnet"].iresnet(192)}[name]()
    else:
        src = io.model_create_by_name(name, 192)
    for p in src.parameters():
        p.data.normal_(0, 0.05)
    path = str(tmp_path / "weight_best.pkl")
    torch.save({"state_dict": src.state_dict()}, path)
    dst = io.model_create_by_name(name, 192)
    res = io.load_weights(dst, path)
    assert not res.missing_keys and not res.unexpected_keys
    sd = dst.state_dict()
    for k, v in src.state_dict().items():
        assert torch.equal(sd[k], v), k
    # full checkpoints (model_checkpoint.pkl: epoch, best_prec, optim) and DataParallel prefixes load too
    torch.save({"epoch": 3, "best_prec": 1.0, "state_dict": {"module." + k: v for k, v in src.state_dict().items()}, "optim": {}}, path)
    assert not io.load_weights(io.model_create_by_name(name, 192), path).missing_keys
