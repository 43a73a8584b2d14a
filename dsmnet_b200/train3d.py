"""Training-mode building blocks of the 3-D stacks (PSMNet hourglass, GC-Net enc-dec).

Custom CUDA here: every convolution, forward and backward (`conv3d.Conv3dFunction`: tcgen05 forward /
dgrad kernels, `dsm_conv3d_wgrad`), BatchNorm3d with batch statistics + ReLU + skip add between them
(`BnActFunction`: the dsm_bn_* streaming kernels on the padded bf16 volumes, forward and backward), the
concat volume (`cost_volume`, fwd+bwd) and the soft-argmin heads (`softargmin`, fwd+bwd).  Frozen (eval-mode)
statistics under autograd and skip tensors that need cropping (odd sizes) have their own kernels
(`BnEvalActFunction`, `CropAddFunction`); there is no stock-PyTorch route.  The inference path
(`psmnet.PSMNetHotPath.aggregate`) stays fully fused into the convolution epilogues; this path exists so
that the 3-D stack can be trained / fine-tuned on the sm_100a kernels.
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


class BnActFunction(torch.autograd.Function):
    """z = act(BatchNorm_batchstats(y) [+ residual]) on padded bf16 volumes (flat storage in, flat storage out).

    forward : dsm_bn_stats -> dsm_bn_finalize_fwd (also the running-statistics update) -> dsm_bn_act_fwd
    backward: dsm_bn_act_bwd_reduce -> dsm_bn_finalize_bwd -> dsm_bn_act_bwd
    Saved for backward: y, z (bf16) and 4*C floats.  relu as in `conv_bn_act`."""

    @staticmethod
    def forward(ctx, ydata, gamma, beta, resdata, geom, relu, eps, momentum, running_mean, running_var, conv_bias):
        B, C, D, H, W = geom
        _lib.require_cuda(ydata, gamma, beta, resdata)
        L, dev = _lib.lib(), ydata.device
        st = _lib.stream_ptr(dev)
        ydata = ydata.contiguous()
        if resdata is not None:
            resdata = resdata.contiguous()
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        stats = torch.empty(4, C, device=dev, dtype=torch.float32)          # scale, shift, mean, rstd
        g32 = None if gamma is None else gamma.detach().float().contiguous()
        b32 = None if beta is None else beta.detach().float().contiguous()
        cb32 = None if conv_bias is None else conv_bias.detach().float().contiguous()
        count = B * D * H * W
        _lib.check(L.dsm_bn_stats(ydata.data_ptr(), B, C, D, H, W, sums.data_ptr(), st), "dsm_bn_stats")
        _lib.check(L.dsm_bn_finalize_fwd(sums.data_ptr(), _lib.ptr(g32), _lib.ptr(b32), _lib.ptr(cb32), C, count,
                                         float(eps), float(momentum), _lib.ptr(running_mean), _lib.ptr(running_var),
                                         stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), st),
                   "dsm_bn_finalize_fwd")
        z = torch.empty_like(ydata)
        _lib.check(L.dsm_bn_act_fwd(ydata.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(), _lib.ptr(resdata), relu,
                                    z.data_ptr(), B, C, D, H, W, st), "dsm_bn_act_fwd")
        ctx.save_for_backward(ydata, z if relu == 1 else None, stats, g32)
        ctx.geom, ctx.relu, ctx.has_res = geom, relu, resdata is not None
        ctx.param_dtype = None if gamma is None else gamma.dtype
        ctx.bias_like = None if conv_bias is None else (conv_bias.shape, conv_bias.dtype)
        return z

    @staticmethod
    def backward(ctx, gz):
        ydata, z, stats, g32 = ctx.saved_tensors
        B, C, D, H, W = ctx.geom
        relu = ctx.relu
        L, dev = _lib.lib(), ydata.device
        st = _lib.stream_ptr(dev)
        gz = gz.contiguous()
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        out = torch.empty(5, C, device=dev, dtype=torch.float32)            # dgamma, dbeta, a, b, c
        _lib.check(L.dsm_bn_act_bwd_reduce(gz.data_ptr(), ydata.data_ptr(), _lib.ptr(z), stats[0].data_ptr(), stats[1].data_ptr(),
                                           relu, sums.data_ptr(), B, C, D, H, W, st), "dsm_bn_act_bwd_reduce")
        _lib.check(L.dsm_bn_finalize_bwd(sums.data_ptr(), _lib.ptr(g32), stats[2].data_ptr(), stats[3].data_ptr(), C, B * D * H * W,
                                         out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), st), "dsm_bn_finalize_bwd")
        dy = torch.empty_like(ydata)
        # the skip tensor's gradient: the masked gradient when the ReLU sits after the add, gz itself otherwise
        need_gres = ctx.has_res and ctx.needs_input_grad[3]
        gres = torch.empty_like(ydata) if (need_gres and relu == 1) else None
        _lib.check(L.dsm_bn_act_bwd(gz.data_ptr(), ydata.data_ptr(), _lib.ptr(z), stats[0].data_ptr(), stats[1].data_ptr(),
                                    out[2].data_ptr(), relu, dy.data_ptr(), _lib.ptr(gres), B, C, D, H, W, st), "dsm_bn_act_bwd")
        if need_gres and relu != 1:
            gres = gz
        dgamma = out[0].to(ctx.param_dtype) if (g32 is not None and ctx.needs_input_grad[1]) else None
        dbeta = out[1].to(ctx.param_dtype) if (ctx.param_dtype is not None and ctx.needs_input_grad[2]) else None
        # a bias in front of batch-statistics BN has an identically zero gradient (sum of dy over a channel is 0)
        gbias = torch.zeros(ctx.bias_like[0], device=dev, dtype=ctx.bias_like[1]) if (ctx.bias_like and ctx.needs_input_grad[10]) else None
        return dy, dgamma, dbeta, gres, None, None, None, None, None, None, gbias


class BnEvalActFunction(torch.autograd.Function):
    """z = act(y*scale + shift [+ residual]) with FROZEN statistics (eval-mode BatchNorm under autograd, e.g. fine-tuning
    with frozen BN): scale = gamma*rstd, shift = beta + (conv_bias - running_mean)*scale.  Same streaming kernels as the
    batch-statistics form: dsm_bn_act_fwd; backward dsm_bn_act_bwd_reduce (sum g, sum g*y) + dsm_bn_act_bwd with
    coef = (scale, 0, 0)."""

    @staticmethod
    def forward(ctx, ydata, gamma, beta, resdata, geom, relu, eps, running_mean, running_var, conv_bias):
        B, C, D, H, W = geom
        _lib.require_cuda(ydata, resdata)
        L, dev = _lib.lib(), ydata.device
        st = _lib.stream_ptr(dev)
        ydata = ydata.contiguous()
        if resdata is not None:
            resdata = resdata.contiguous()
        rstd = torch.rsqrt(running_var.detach().float() + eps)
        g32 = torch.ones(C, device=dev) if gamma is None else gamma.detach().float()
        b32 = torch.zeros(C, device=dev) if beta is None else beta.detach().float()
        mean = running_mean.detach().float()
        if conv_bias is not None:
            mean = mean - conv_bias.detach().float()
        scale = (g32 * rstd).contiguous()
        shift = (b32 - mean * scale).contiguous()
        z = torch.empty_like(ydata)
        _lib.check(L.dsm_bn_act_fwd(ydata.data_ptr(), scale.data_ptr(), shift.data_ptr(), _lib.ptr(resdata), relu,
                                    z.data_ptr(), B, C, D, H, W, st), "dsm_bn_act_fwd")
        ctx.save_for_backward(ydata, z if relu == 1 else None, scale, shift, rstd, mean)
        ctx.geom, ctx.relu, ctx.has_res = geom, relu, resdata is not None
        ctx.param_dtype = None if gamma is None else gamma.dtype
        ctx.bias_like = None if conv_bias is None else (conv_bias.shape, conv_bias.dtype)
        return z

    @staticmethod
    def backward(ctx, gz):
        ydata, z, scale, shift, rstd, mean = ctx.saved_tensors
        B, C, D, H, W = ctx.geom
        relu = ctx.relu
        L, dev = _lib.lib(), ydata.device
        st = _lib.stream_ptr(dev)
        gz = gz.contiguous()
        sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
        _lib.check(L.dsm_bn_act_bwd_reduce(gz.data_ptr(), ydata.data_ptr(), _lib.ptr(z), scale.data_ptr(), shift.data_ptr(),
                                           relu, sums.data_ptr(), B, C, D, H, W, st), "dsm_bn_act_bwd_reduce")
        coef = torch.cat((scale, torch.zeros(2 * C, device=dev))).contiguous()           # dy = scale * g
        dy = torch.empty_like(ydata)
        need_gres = ctx.has_res and ctx.needs_input_grad[3]
        gres = torch.empty_like(ydata) if (need_gres and relu == 1) else None
        _lib.check(L.dsm_bn_act_bwd(gz.data_ptr(), ydata.data_ptr(), _lib.ptr(z), scale.data_ptr(), shift.data_ptr(),
                                    coef.data_ptr(), relu, dy.data_ptr(), _lib.ptr(gres), B, C, D, H, W, st), "dsm_bn_act_bwd")
        if need_gres and relu != 1:
            gres = gz
        sg, sgy = sums[:C].float(), sums[C:].float()
        dgamma = dbeta = gbias = None
        if ctx.param_dtype is not None:
            if ctx.needs_input_grad[1]:
                dgamma = (rstd * (sgy - mean * sg)).to(ctx.param_dtype)        # sum g * (y + bias - running_mean) * rstd
            if ctx.needs_input_grad[2]:
                dbeta = sg.to(ctx.param_dtype)
        if ctx.bias_like and ctx.needs_input_grad[9]:
            gbias = (scale * sg).reshape(ctx.bias_like[0]).to(ctx.bias_like[1])
        return dy, dgamma, dbeta, gres, None, None, None, None, None, gbias


class CropAddFunction(torch.autograd.Function):
    """z = act(crop(full) + residual): the reference's crop-to-min skip add (myadd_3d, stackhourglass.py:10-20; myAdd3d,
    util_fun.py:41-51) when the deconv output is larger than the skip tensor (odd sizes).  geom = (B, C, natural extent,
    cropped extent); relu: ReLU after the add (PSMNet) or none (GC-Net activates before the add)."""

    @staticmethod
    def forward(ctx, fulldata, resdata, geom, relu):
        B, C, nat, od = geom
        _lib.require_cuda(fulldata, resdata)
        dev = fulldata.device
        z = torch.empty(B * (od[0] + 2) * (od[1] + 2) * (od[2] + 2) * C, device=dev, dtype=torch.bfloat16)
        _lib.check(_lib.lib().dsm_crop_add_fwd(fulldata.contiguous().data_ptr(), resdata.contiguous().data_ptr(), z.data_ptr(),
                                               B, C, *nat, *od, int(relu), _lib.stream_ptr(dev)), "dsm_crop_add_fwd")
        ctx.save_for_backward(z if relu else None)
        ctx.geom, ctx.relu = geom, int(relu)
        return z

    @staticmethod
    def backward(ctx, gz):
        (z,) = ctx.saved_tensors
        B, C, nat, od = ctx.geom
        dev = gz.device
        gz = gz.contiguous()
        gfull = torch.empty(B * (nat[0] + 2) * (nat[1] + 2) * (nat[2] + 2) * C, device=dev, dtype=torch.bfloat16)
        gres = PaddedVolume.empty_zero_rim(B, C, *od, dev).data if ctx.needs_input_grad[1] else None
        _lib.check(_lib.lib().dsm_crop_add_bwd(gz.data_ptr(), _lib.ptr(z), gfull.data_ptr(), _lib.ptr(gres), B, C, *nat, *od,
                                               ctx.relu, _lib.stream_ptr(dev)), "dsm_crop_add_bwd")
        return gfull, gres, None, None


def bn_act(y: PaddedVolume, bn: nn.BatchNorm3d, relu: int = 0, residual: Optional[PaddedVolume] = None,
           conv_bias: Optional[torch.Tensor] = None) -> PaddedVolume:
    """BatchNorm3d + act + skip on the fused kernels.  Training mode: batch statistics, running statistics updated as
    nn.BatchNorm3d does.  Eval mode (frozen statistics under autograd): `BnEvalActFunction`."""
    if residual is not None and (residual.B, residual.C, residual.D, residual.H, residual.W) != (y.B, y.C, y.D, y.H, y.W):
        raise _lib.DsmError("bn_act: the skip tensor must have the geometry of y")
    geom = (y.B, y.C, y.D, y.H, y.W)
    if not bn.training and bn.track_running_stats:
        z = BnEvalActFunction.apply(y.data, bn.weight, bn.bias, None if residual is None else residual.data, geom, int(relu),
                                    bn.eps, bn.running_mean, bn.running_var, conv_bias)
        return PaddedVolume(z, *geom)
    momentum = bn.momentum
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    track = bn.track_running_stats
    z = BnActFunction.apply(y.data, bn.weight, bn.bias, None if residual is None else residual.data,
                            geom, int(relu), bn.eps, momentum if momentum is not None else 0.1,
                            bn.running_mean if track else None, bn.running_var if track else None, conv_bias)
    return PaddedVolume(z, *geom)


def crop_add(full: PaddedVolume, residual: PaddedVolume, relu: bool) -> PaddedVolume:
    od = (min(full.D, residual.D), min(full.H, residual.H), min(full.W, residual.W))
    if od != (residual.D, residual.H, residual.W) or full.C != residual.C or full.B != residual.B:
        raise _lib.DsmError("crop_add: the skip tensor must not exceed the deconv output (crop-to-min of the reference's graphs)")
    z = CropAddFunction.apply(full.data, residual.data, (full.B, full.C, (full.D, full.H, full.W), od), bool(relu))
    return PaddedVolume(z, full.B, full.C, *od)


def conv_bn_act(x: PaddedVolume, conv: nn.Module, bn: Optional[nn.BatchNorm3d], relu: int = 0,
                residual: Optional[PaddedVolume] = None) -> PaddedVolume:
    """conv (+bias) -> BatchNorm3d (batch statistics when bn.training, frozen otherwise) -> [ReLU] -> [+ residual,
    crop-to-min] -> [ReLU].  relu: 0 none, 1 after the residual add (PSMNet), 2 before it (GC-Net).
    Every step is one of the library's kernels; there is no stock-PyTorch route.  With a skip tensor smaller than the
    deconv output (odd sizes) BatchNorm sees the full output, as in the reference, and the crop happens at the add."""
    if bn is None:
        raise _lib.DsmError("conv_bn_act: every >= 32-channel 3-D layer of PSMNet / GC-Net carries a BatchNorm3d; "
                            "the single-channel layers go through conv_c1")
    transposed = isinstance(conv, nn.ConvTranspose3d)
    stride = conv.stride[0]
    nat = conv_out_dims(x.D, x.H, x.W, stride, transposed)
    y = conv3d_train(x, conv.weight, stride, transposed, nat)
    if y.C not in (32, 64, 128):
        raise _lib.DsmError("conv_bn_act: %d output channels are not supported by the BatchNorm kernels" % y.C)
    if residual is None or (residual.D, residual.H, residual.W) == nat:
        return bn_act(y, bn, relu, residual, conv.bias)
    z = bn_act(y, bn, 2 if relu == 2 else 0, None, conv.bias)
    return crop_add(z, residual, relu == 1)


_c1_ws = {}


class _ConvC1Function(torch.autograd.Function):
    """Single-output-channel layers (PSMNet classif*.2: Conv3d 32->1, GC-Net l37: ConvTranspose3d 32->1):
    fp32 output [B, Do, Ho, Wo].  Forward on the tcgen05 kernels; backward is dsm_conv3d_c1_bwd, one CUDA-core pass
    that produces the input gradient and the weight gradient from the 27 gathered gy values of each voxel."""

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
        w = weight.detach().float()
        w27 = (w[:, 0] if transposed else w[0]).reshape(C, 27).contiguous()
        gxv = PaddedVolume.empty_zero_rim(B, C, D, H, W, dev)
        dw = torch.empty(C, 27, device=dev, dtype=torch.float32)
        L = _lib.lib()
        ws = _c1_ws.get(str(dev))
        if ws is None:
            ws = _c1_ws[str(dev)] = torch.empty(L.dsm_conv3d_c1_bwd_workspace_bytes() // 4, device=dev, dtype=torch.float32)
        _lib.check(L.dsm_conv3d_c1_bwd(xdata.data_ptr(), gy.data_ptr(), w27.data_ptr(), gxv.data.data_ptr(), dw.data_ptr(),
                                       B, D, H, W, Do, Ho, Wo, int(transposed), ws.data_ptr(), ws.numel() * 4,
                                       _lib.stream_ptr(dev)), "dsm_conv3d_c1_bwd")
        gx = gxv.data if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            gw = (dw.view(C, 1, 3, 3, 3) if transposed else dw.view(1, C, 3, 3, 3)).to(weight.dtype)
        gb = gy.sum().reshape(1) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


def conv_c1(x: PaddedVolume, conv: nn.Module) -> torch.Tensor:
    """Differentiable single-channel conv / transposed conv -> fp32 [B, Do, Ho, Wo]."""
    if x.C != 32:
        raise _lib.DsmError("conv_c1: the single-channel layers of PSMNet / GC-Net take 32 input channels")
    transposed = isinstance(conv, nn.ConvTranspose3d)
    return _ConvC1Function.apply(x.data, conv.weight, conv.bias, (x.B, x.C, x.D, x.H, x.W, transposed))
