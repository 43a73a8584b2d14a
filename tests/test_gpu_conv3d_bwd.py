"""GPU (B200): backward of the 3-D convolution layers (training path).  dgrad runs on the forward
tcgen05 kernels with transformed weights, wgrad on dsm_conv3d_wgrad; both are compared with torch
autograd of the fp32 definition on bf16-rounded operands (the oracle's conv3d_block)."""
import pytest
import torch
import torch.nn.functional as F

import oracle.ops as O

pytestmark = pytest.mark.gpu


def l2rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


CASES = [
    # cin, cout, D, H, W, stride, transposed
    (32, 32, 4, 6, 20, 1, False),
    (64, 32, 5, 7, 19, 1, False),
    (64, 64, 9, 6, 70, 1, False),        # W > 64: two anchor tiles per row; D > 8: two bands
    (32, 64, 8, 12, 40, 2, False),
    (64, 64, 7, 11, 37, 2, False),       # odd sizes: the stride-2 dgrad is a cropped transposed conv
    (64, 32, 3, 5, 10, 2, True),
    (64, 64, 4, 6, 33, 2, True),
    (128, 64, 3, 4, 9, 2, True),
    (64, 128, 6, 8, 18, 2, False),
    (128, 128, 3, 4, 9, 1, False),
]


@pytest.mark.parametrize("cin,cout,D,H,W,stride,transposed", CASES)
def test_conv3d_backward_vs_autograd(cin, cout, D, H, W, stride, transposed):
    from dsmnet_b200.conv3d import conv3d_train, conv_timeouts
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(0)
    x = torch.randn(2, cin, D, H, W).to(torch.bfloat16).float()
    w = (torch.randn(cin, cout, 3, 3, 3) if transposed else torch.randn(cout, cin, 3, 3, 3)) * 0.05
    w = w.to(torch.bfloat16).float()
    # reference: fp32 conv of the same bf16-representable operands
    xr = x.clone().requires_grad_(); wr = w.clone().requires_grad_()
    yr = O.conv3d_block(xr, wr, None, None, stride, transposed)
    gy = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(gy)
    # CUDA path
    xv = PaddedVolume.from_ncdhw(x.cuda())
    xd = xv.data.clone().requires_grad_()
    wc = w.cuda().requires_grad_()
    y = conv3d_train(PaddedVolume(xd, *xv.shape5[:1], cin, D, H, W), wc, stride, transposed)
    assert (y.D, y.H, y.W) == tuple(yr.shape[2:])
    assert l2rel(y.to_ncdhw().cpu(), yr.detach()) < 1e-2                         # bf16 output rounding
    gyv = PaddedVolume.from_ncdhw(gy.cuda())
    y.data.backward(gyv.data)
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    gx = PaddedVolume(xd.grad, 2, cin, D, H, W).to_ncdhw().cpu()
    assert l2rel(gx, xr.grad) < 1e-2                                             # dgrad (bf16 output rounding)
    assert l2rel(wc.grad.cpu(), wr.grad) < 2e-3                                  # wgrad (fp32 accumulate and output)
    # rim of the input gradient stays zero (the layout's invariant)
    g6 = xd.grad.view(2, D + 2, H + 2, W + 2, cin)
    assert float(g6[:, 0].abs().max()) == 0.0 and float(g6[:, :, :, 0].abs().max()) == 0.0
