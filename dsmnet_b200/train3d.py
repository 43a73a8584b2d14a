"""Training-mode building blocks of the 3-D stacks (PSMNet hourglass, GC-Net enc-dec).

What is custom CUDA here: every convolution, forward and backward (`conv3d.Conv3dFunction`: tcgen05
forward / dgrad kernels, `dsm_conv3d_wgrad`), the concat volume (`cost_volume`, fwd+bwd) and the
soft-argmin (`softargmin`, fwd+bwd).  BatchNorm with batch statistics, ReLU, the skip adds and the
trilinear upsample of the training graph are stock PyTorch ops on views of the padded volumes — the
reference's are stock modules too (submodule.py:16-19, stackhourglass.py:43-62,152-166).  The
inference path (`psmnet.PSMNetHotPath.aggregate`) stays fully fused; this path exists so that the
3-D stack can be trained / fine-tuned without leaving the sm_100a kernels for the convolutions.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .conv3d import FusedConv3d, _dgrad_layer, conv3d_train, conv3d_wgrad, conv_out_dims
from .volume_layout import PaddedVolume


def interior(v: PaddedVolume) -> torch.Tensor:
    """[B, D, H, W, C] view of the voxels inside the zero rim (shares storage, differentiable)."""
    return v.data.view(v.B, v.D + 2, v.H + 2, v.W + 2, v.C)[:, 1:-1, 1:-1, 1:-1, :]


def from_interior(t: torch.Tensor) -> PaddedVolume:
    """[B, D, H, W, C] (any float dtype) -> PaddedVolume (bf16, zero rim); differentiable."""
    B, D, H, W, C = t.shape
    p = F.pad(t.to(torch.bfloat16), (0, 0, 1, 1, 1, 1, 1, 1))
    return PaddedVolume(p.reshape(-1), B, C, D, H, W)


def volume_from_ncdhw(x: torch.Tensor) -> PaddedVolume:
    """NCDHW fp32 -> PaddedVolume, differentiable (the inference path uses the dsm_pack_ndhwc kernel)."""
    return from_interior(x.permute(0, 2, 3, 4, 1))


def conv_bn_act(x: PaddedVolume, conv: nn.Module, bn: Optional[nn.BatchNorm3d], relu: int = 0,
                residual: Optional[PaddedVolume] = None) -> PaddedVolume:
    """conv (+bias) -> BatchNorm3d (batch statistics when bn.training) -> [ReLU] -> [+ residual, crop-to-min] -> [ReLU].
    relu: 0 none, 1 after the residual add (PSMNet), 2 before it (GC-Net)."""
    transposed = isinstance(conv, nn.ConvTranspose3d)
    stride = conv.stride[0]
    nat = conv_out_dims(x.D, x.H, x.W, stride, transposed)
    od = nat if residual is None else (min(nat[0], residual.D), min(nat[1], residual.H), min(nat[2], residual.W))
    y = conv3d_train(x, conv.weight, stride, transposed, od)
    z = interior(y).float()
    if conv.bias is not None:
        z = z + conv.bias
    if bn is not None:
        z = F.batch_norm(z.permute(0, 4, 1, 2, 3), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                         bn.training, bn.momentum if bn.momentum is not None else 0.1, bn.eps).permute(0, 2, 3, 4, 1)
    if relu == 2:
        z = F.relu(z)
    if residual is not None:
        z = z + interior(residual)[:, :od[0], :od[1], :od[2], :].float()
    if relu == 1:
        z = F.relu(z)
    return from_interior(z)


class _ConvC1Function(torch.autograd.Function):
    """Single-output-channel layers (PSMNet classif*.2: Conv3d 32->1, GC-Net l37: ConvTranspose3d 32->1):
    fp32 output [B, Do, Ho, Wo].  Backward pads the one gradient channel to 32 so that the same dgrad /
    wgrad kernels apply (31 zero channels: 32x redundant MACs on a layer that is 0.3 % of the flops)."""

    @staticmethod
    def forward(ctx, xdata, weight, bias, geom):
        B, C, D, H, W, transposed = geom
        x = PaddedVolume(xdata, B, C, D, H, W)
        layer = FusedConv3d(weight, None, bias, 2 if transposed else 1, transposed, 0, xdata.device)
        y = layer(x)
        ctx.save_for_backward(xdata, weight)
        ctx.geom, ctx.has_bias = geom, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        xdata, weight = ctx.saved_tensors
        B, C, D, H, W, transposed = ctx.geom
        dev = xdata.device
        gy = gy.contiguous().float()
        Do, Ho, Wo = gy.shape[-3:]
        g5 = torch.zeros(B, Do + 2, Ho + 2, Wo + 2, 32, device=dev, dtype=torch.bfloat16)
        g5[:, 1:-1, 1:-1, 1:-1, 0] = gy.to(torch.bfloat16)
        g = PaddedVolume(g5.reshape(-1), B, 32, Do, Ho, Wo)
        x = PaddedVolume(xdata, B, C, D, H, W)
        w = weight.detach()
        if transposed:                              # weight [C][1] -> [C][32]
            w32 = torch.zeros(C, 32, 3, 3, 3, device=dev, dtype=w.dtype); w32[:, :1] = w
        else:                                       # weight [1][C] -> [32][C]
            w32 = torch.zeros(32, C, 3, 3, 3, device=dev, dtype=w.dtype); w32[:1] = w
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _dgrad_layer(w32, 2 if transposed else 1, transposed, dev)(g, PaddedVolume.empty(B, C, D, H, W, dev)).data
        if ctx.needs_input_grad[1]:
            gw = conv3d_wgrad(x, g, 2, C, 1) if transposed else conv3d_wgrad(g, x, 1, 1, C)
            gw = gw.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum().reshape(1)
        return gx, gw, gb, None


def conv_c1(x: PaddedVolume, conv: nn.Module) -> torch.Tensor:
    """Differentiable single-channel conv / transposed conv -> fp32 [B, Do, Ho, Wo]."""
    if x.C != 32:
        raise _lib.DsmError("conv_c1: the single-channel layers of PSMNet / GC-Net take 32 input channels")
    transposed = isinstance(conv, nn.ConvTranspose3d)
    return _ConvC1Function.apply(x.data, conv.weight, conv.bias, (x.B, x.C, x.D, x.H, x.W, transposed))
