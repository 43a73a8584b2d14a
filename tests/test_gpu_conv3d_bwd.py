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


def _wgrad_both_modes(cin, cout, B, D, H, W):
    """dW of a stride-1 conv from the tcgen05 kernel (mode 0) and from the warp-level kernel (mode 1)"""
    from dsmnet_b200 import _lib
    from dsmnet_b200.conv3d import conv3d_wgrad
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(3)
    x = PaddedVolume.from_ncdhw(torch.randn(B, cin, D, H, W, device="cuda"))
    gy = PaddedVolume.from_ncdhw(torch.randn(B, cout, D, H, W, device="cuda"))
    L = _lib.lib()
    out = []
    for mode in (0, 1):
        prev = L.dsm_debug_wgrad_mode(mode)
        try:
            out.append(conv3d_wgrad(gy, x, 1, cout, cin).clone())
        finally:
            L.dsm_debug_wgrad_mode(prev)
    torch.cuda.synchronize()
    assert L.dsm_debug_wgrad_timeouts() == 0
    return out, x, gy


@pytest.mark.parametrize("cin,cout,B,D,H,W", [
    (32, 32, 1, 12, 24, 78),       # hourglass quarter resolution
    (32, 32, 2, 1, 3, 5),          # one plane, one chunk
    (64, 32, 1, 7, 14, 141),       # partner channel blocks; plane size not a multiple of the 128-position step
    (64, 64, 3, 5, 9, 33),
    (32, 128, 1, 4, 6, 50),
])
def test_wgrad_tcgen05_matches_warp_level_kernel(cin, cout, B, D, H, W):
    (tcg, leg), x, gy = _wgrad_both_modes(cin, cout, B, D, H, W)
    assert l2rel(tcg, leg) < 1e-5                                                 # same bf16 products, fp32 sums in another order
    # and both against autograd of the fp32 definition on the same bf16 operands
    xr = x.to_ncdhw().requires_grad_(False)
    wr = torch.zeros(cout, cin, 3, 3, 3, device="cuda", requires_grad=True)
    F.conv3d(xr, wr, padding=1).backward(gy.to_ncdhw())
    assert l2rel(tcg, wr.grad) < 2e-3


def test_wgrad_tcgen05_full_size():
    """BASELINE size (32 -> 32 at 48 x 96 x 312): tcgen05 wgrad against the warp-level kernel"""
    (tcg, leg), _, _ = _wgrad_both_modes(32, 32, 1, 48, 96, 312)
    assert l2rel(tcg, leg) < 1e-5


@pytest.mark.parametrize("transposed,B,D,H,W", [
    (False, 1, 4, 6, 20), (False, 2, 3, 5, 70), (False, 1, 2, 3, 129),
    (True, 1, 3, 4, 9), (True, 2, 2, 5, 66), (True, 1, 5, 3, 64),
])
def test_conv_c1_backward_vs_autograd(transposed, B, D, H, W):
    """32 -> 1 layers (PSMNet classif*.2, GC-Net l37): dsm_conv3d_c1_bwd against autograd of the fp32 definition"""
    import torch.nn as nn
    from dsmnet_b200 import train3d as T
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(4)
    conv = (nn.ConvTranspose3d(32, 1, 3, stride=2, padding=1, output_padding=1) if transposed
            else nn.Conv3d(32, 1, 3, padding=1)).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    x = torch.randn(B, 32, D, H, W, device="cuda").to(torch.bfloat16).float()
    # reference in fp64 (cuDNN's fp32 convolutions may run in TF32)
    xr = x.double().requires_grad_()
    conv.double()
    yr = conv(xr)
    gy = torch.randn_like(yr).float().double()
    yr.backward(gy)
    ref_gw, ref_gb = conv.weight.grad.float().clone(), conv.bias.grad.float().clone()
    conv.zero_grad(); conv.float()
    yr, gy = yr.float(), gy.float()
    xv = PaddedVolume.from_ncdhw(x)
    xd = xv.data.clone().requires_grad_()
    y = T.conv_c1(PaddedVolume(xd, B, 32, D, H, W), conv)
    assert tuple(y.shape) == (B,) + tuple(yr.shape[2:])
    assert l2rel(y, yr.detach().squeeze(1)) < 1e-2
    y.backward(gy.squeeze(1))
    gx = PaddedVolume(xd.grad, B, 32, D, H, W).to_ncdhw()
    assert l2rel(gx, xr.grad.float()) < 5e-3                  # bf16 output rounding
    assert l2rel(conv.weight.grad, ref_gw) < 1e-4             # fp32 throughout
    assert l2rel(conv.bias.grad, ref_gb) < 1e-5
    g6 = xd.grad.view(B, D + 2, H + 2, W + 2, 32)
    assert float(g6[:, 0].abs().max()) == 0.0 and float(g6[:, :, :, -1].abs().max()) == 0.0


@pytest.mark.parametrize("ca,cb,B,Da,Ha,Wa,Dp,Hp,Wp", [
    (64, 32, 1, 6, 12, 39, 12, 24, 78),      # conv1-like (32 -> 64, stride 2): anchor = gy (coarse), partner = x (fine)
    (64, 64, 2, 4, 6, 19, 7, 11, 37),        # odd fine sizes
    (32, 32, 1, 3, 5, 10, 6, 10, 20),
    (64, 128, 1, 3, 4, 9, 6, 8, 18),
    (32, 32, 1, 3, 4, 5, 5, 7, 9),           # transposed layer with a cropped output (partner smaller than 2x)
])
def test_wgrad_stride2_tcgen05_matches_warp_level_kernel(ca, cb, B, Da, Ha, Wa, Dp, Hp, Wp):
    """stride-2 / transposed layers: parity gather + tcgen05 kernel (mode 0) against the warp-level kernel (mode 1)"""
    from dsmnet_b200 import _lib
    from dsmnet_b200.conv3d import conv3d_wgrad
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(5)
    anchor = PaddedVolume.from_ncdhw(torch.randn(B, ca, Da, Ha, Wa, device="cuda"))
    partner = PaddedVolume.from_ncdhw(torch.randn(B, cb, Dp, Hp, Wp, device="cuda"))
    L = _lib.lib()
    out = []
    for mode in (0, 1):
        prev = L.dsm_debug_wgrad_mode(mode)
        try:
            out.append(conv3d_wgrad(anchor, partner, 2, ca, cb).clone())
        finally:
            L.dsm_debug_wgrad_mode(prev)
    torch.cuda.synchronize()
    assert L.dsm_debug_wgrad_timeouts() == 0
    assert l2rel(out[0], out[1]) < 1e-5
