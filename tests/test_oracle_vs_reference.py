"""CPU, build container only: oracle/ops.py against the reference executed live through the shim.
Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import pytest
import torch

import oracle.ops as O
import oracle.refshim as R
from conftest import rel_err

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")


@pytest.mark.parametrize("D,s,k", [(7, 1, 1), (9, 2, 3), (30, 1, 1)])
def test_corr1d_live(D, s, k):
    torch.manual_seed(0)
    fL = torch.relu(torch.randn(1, 16, 5, 24)); fR = torch.relu(torch.randn(1, 16, 5, 24))
    assert torch.equal(O.corr1d(fL, fR, D, s, k), R.corr1d(fL, fR, D, s, k))


def test_volumes_live():
    torch.manual_seed(1)
    fL = torch.randn(2, 4, 3, 10); fR = torch.randn(2, 4, 3, 10)
    assert torch.equal(O.concat_volume(fL, fR, 6, "psm"), R.psm_volume(fL, fR, 24))
    assert torch.equal(O.concat_volume(fL, fR, 6, "gc"), R.gc_volume(fL, fR, 6))
    assert torch.equal(O.concat_volume(fL, fR, 6, "gc_right"), R.gc_volume(fL, fR, 6, True))


def test_imwrap_live():
    torch.manual_seed(2)
    src = torch.rand(2, 3, 20, 30); disp = torch.rand(2, 1, 10, 15) * 3
    for kw in (dict(), dict(fliplr=True), dict(LeftTop=(4, 2))):
        out, delt = R.imwrap(src, disp, **kw)
        assert torch.equal(out, O.imwrap(src, disp, delt=delt, **kw))


def test_psmnet_live():
    torch.manual_seed(3)
    fL = torch.randn(1, 32, 8, 12); fR = torch.randn(1, 32, 8, 12)
    cost = R.psm_volume(fL, fR, 16)
    params = O.psmnet_random_params(seed=3, calibrate_on=cost)
    net = R.make_psmnet(16, 0)
    net.load_state_dict(params, strict=False)
    ref = R.psmnet_forward_from_features(net, fL, fR, 16, 32, 48)
    mine = O.psmnet_hotpath(params, fL, fR, 16, (32, 48))
    for a, b in zip(mine, ref):
        assert (a - b).abs().max() < 2e-3
