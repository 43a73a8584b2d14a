"""Fused training-mode BatchNorm3d + ReLU + skip add (dsm_bn_*) against stock PyTorch ops on the same bf16 inputs.

Reference semantics: nn.BatchNorm3d under model.train() followed by F.relu / the skip adds, models/psmnet/submodule.py:16-19,
models/psmnet/stackhourglass.py:43-62, models/util_conv.py:160-178.  Tolerances: the outputs are bf16 (relative 2^-8 per
element), the per-channel gradients fp32 sums of bf16 products.
"""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _padded(t):
    """[B, D, H, W, C] fp32 -> flat padded bf16 storage"""
    return F.pad(t.to(torch.bfloat16), (0, 0, 1, 1, 1, 1, 1, 1)).reshape(-1).contiguous()


def _interior(flat, B, C, D, H, W):
    return flat.view(B, D + 2, H + 2, W + 2, C)[:, 1:-1, 1:-1, 1:-1, :]


def _reference(y, res, bn, relu):
    """stock ops, fp32, on the bf16-rounded inputs; y/res: [B, D, H, W, C] fp32 leaf tensors"""
    z = bn(y.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    if relu == 2:
        z = F.relu(z)
    if res is not None:
        z = z + res
    if relu == 1:
        z = F.relu(z)
    return z


@pytest.mark.parametrize("C,dims,relu,with_res,B", [
    (32, (6, 10, 21), 1, False, 1),
    (32, (5, 9, 20), 1, True, 2),
    (32, (4, 8, 17), 0, True, 1),
    (64, (4, 7, 13), 2, True, 1),
    (64, (3, 6, 11), 2, False, 2),
    (128, (3, 5, 9), 2, True, 1),
    (64, (4, 6, 10), 0, False, 1),
])
def test_bn_act_matches_torch(C, dims, relu, with_res, B):
    from dsmnet_b200 import train3d as T
    from dsmnet_b200.volume_layout import PaddedVolume
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(C + relu * 7 + B)
    D, H, W = dims
    y0 = (torch.randn(B, D, H, W, C, generator=g) * 1.7 + 0.4).to(torch.bfloat16).float().to(dev)
    r0 = torch.randn(B, D, H, W, C, generator=g).to(torch.bfloat16).float().to(dev) if with_res else None
    gz0 = torch.randn(B, D, H, W, C, generator=g).to(torch.bfloat16).float().to(dev)
    bias = torch.randn(C, generator=g).to(dev)

    bn_ref = nn.BatchNorm3d(C).to(dev).train()
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(C, generator=g) + 0.5); bn_ref.bias.copy_(torch.randn(C, generator=g) * 0.3)
    bn_new = nn.BatchNorm3d(C).to(dev).train()
    bn_new.load_state_dict(bn_ref.state_dict())

    # reference (the conv bias is added in front of BN: it must only move the running mean)
    yr = y0.clone().requires_grad_(True)
    rr = r0.clone().requires_grad_(True) if with_res else None
    zr = _reference(yr + bias.view(1, 1, 1, 1, C), rr, bn_ref, relu)
    zr.backward(gz0)

    yn = _padded(y0).requires_grad_(True)
    rn = _padded(r0).requires_grad_(True) if with_res else None
    out = T.bn_act(PaddedVolume(yn, B, C, D, H, W), bn_new, relu,
                   PaddedVolume(rn, B, C, D, H, W) if with_res else None, conv_bias=bias)
    full = out.data.view(B, D + 2, H + 2, W + 2, C).float()
    zi = _interior(out.data, B, C, D, H, W).float()
    # zero rim
    assert full.abs().sum().item() == pytest.approx(zi.abs().sum().item(), rel=1e-6)
    assert torch.allclose(zi, zr.detach(), rtol=1e-2, atol=1e-2), (zi - zr).abs().max().item()
    # running statistics as nn.BatchNorm3d updates them
    assert torch.allclose(bn_new.running_mean, bn_ref.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn_new.running_var, bn_ref.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn_new.num_batches_tracked) == int(bn_ref.num_batches_tracked) == 1

    out.data.backward(_padded(gz0))
    dy = _interior(yn.grad, B, C, D, H, W).float()
    assert yn.grad.view(B, D + 2, H + 2, W + 2, C).float().abs().sum().item() == pytest.approx(dy.abs().sum().item(), rel=1e-6)
    # ReLU masks of elements whose pre-activation is within bf16 rounding of 0 may differ: compare in aggregate
    err = (dy - yr.grad).abs()
    scale = yr.grad.abs().mean().item()
    assert err.mean().item() < 0.01 * scale + 1e-6, (err.mean().item(), scale)
    assert (err > 0.05 * scale + 0.02 * yr.grad.abs()).float().mean().item() < 5e-3
    assert torch.allclose(bn_new.weight.grad, bn_ref.weight.grad, rtol=2e-2, atol=2e-2 * bn_ref.weight.grad.abs().max().item())
    assert torch.allclose(bn_new.bias.grad, bn_ref.bias.grad, rtol=2e-2, atol=2e-2 * bn_ref.bias.grad.abs().max().item())
    if with_res:
        gr = _interior(rn.grad, B, C, D, H, W).float()
        e = (gr - rr.grad).abs()
        assert (e > 1e-6).float().mean().item() < 5e-3


def test_bn_act_rejects_cpu_and_bad_channels():
    from dsmnet_b200 import _lib, train3d as T
    from dsmnet_b200.volume_layout import PaddedVolume
    bn = nn.BatchNorm3d(32).train()
    y = torch.zeros(1 * 4 * 4 * 4 * 32, dtype=torch.bfloat16)
    with pytest.raises(_lib.DsmError):
        T.bn_act(PaddedVolume(y, 1, 32, 2, 2, 2), bn)
    dev = torch.device("cuda")
    bn48 = nn.BatchNorm3d(48).to(dev).train()
    y48 = torch.zeros(1 * 4 * 4 * 4 * 48, dtype=torch.bfloat16, device=dev)
    with pytest.raises(_lib.DsmError):
        T.bn_act(PaddedVolume(y48, 1, 48, 2, 2, 2), bn48)
