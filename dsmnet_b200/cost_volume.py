"""Concatenation cost volume (PSMNet / GC-Net) on the sm_100a kernel.

The reference builds the volume inline with Python loops (models/psmnet/stackhourglass.py:124-133,
models/gcnet.py:131-135 and :156-164); ``concat_volume(fL, fR, D, mode)`` is the function those
loops become.  ``mode``: "psm", "gc" or "gc_right".  The default output is the reference's
NCDHW fp32 tensor (bit-exact copy semantics, differentiable); ``padded_bf16=True`` emits the
3-D stack's own layout (bf16, [B][D+2][H+2][W+2][2C], zero rim) as a ``PaddedVolume``.
"""
from __future__ import annotations

import torch

from . import _lib
from .volume_layout import PaddedVolume


class ConcatVolumeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fL, fR, D, mode):
        _lib.require_cuda(fL, fR)
        fL = fL.contiguous().float(); fR = fR.contiguous().float()
        B, C, H, W = fL.shape
        out = torch.empty(B, 2 * C, D, H, W, device=fL.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_concat_volume_fwd(fL.data_ptr(), fR.data_ptr(), out.data_ptr(), B, C, D, H, W, mode,
                                                    _lib.DSM_F32, _lib.DSM_NCDHW, _lib.stream_ptr(fL.device)),
                   "dsm_concat_volume_fwd")
        ctx.shape, ctx.D, ctx.mode = (B, C, H, W), D, mode
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous().float()
        gL = torch.empty(B, C, H, W, device=g.device, dtype=torch.float32)
        gR = torch.empty_like(gL)
        _lib.check(_lib.lib().dsm_concat_volume_bwd(g.data_ptr(), gL.data_ptr(), gR.data_ptr(), B, C, ctx.D, H, W, ctx.mode,
                                                    _lib.DSM_F32, _lib.DSM_NCDHW, _lib.stream_ptr(g.device)),
                   "dsm_concat_volume_bwd")
        return gL, gR, None, None


class ConcatVolumePaddedFunction(torch.autograd.Function):
    """The volume in the 3-D stack's layout (flat storage of a PaddedVolume, bf16, zero rim), differentiable:
    backward sums the padded bf16 gradient over d straight into the two NCHW fp32 feature gradients."""

    @staticmethod
    def forward(ctx, fL, fR, D, mode):
        _lib.require_cuda(fL, fR)
        fL = fL.contiguous().float(); fR = fR.contiguous().float()
        B, C, H, W = fL.shape
        vol = PaddedVolume.empty(B, 2 * C, D, H, W, fL.device, zero_rim=False)
        _lib.check(_lib.lib().dsm_concat_volume_fwd(fL.data_ptr(), fR.data_ptr(), vol.data.data_ptr(), B, C, D, H, W, mode,
                                                    _lib.DSM_BF16, _lib.DSM_NDHWC_PADDED, _lib.stream_ptr(fL.device)),
                   "dsm_concat_volume_fwd")
        ctx.shape, ctx.D, ctx.mode = (B, C, H, W), D, mode
        return vol.data

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous()
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        gL = torch.empty(B, C, H, W, device=g.device, dtype=torch.float32)
        gR = torch.empty_like(gL)
        _lib.check(_lib.lib().dsm_concat_volume_bwd(g.data_ptr(), gL.data_ptr(), gR.data_ptr(), B, C, ctx.D, H, W, ctx.mode,
                                                    _lib.DSM_BF16, _lib.DSM_NDHWC_PADDED, _lib.stream_ptr(g.device)),
                   "dsm_concat_volume_bwd")
        return gL, gR, None, None


def concat_volume_padded(fL, fR, D, mode="psm") -> PaddedVolume:
    """Differentiable padded-bf16 volume (training path of the 3-D stacks)."""
    if fL.shape != fR.shape or fL.dim() != 4:
        raise _lib.DsmError("concat_volume expects two NCHW tensors of equal shape")
    B, C, H, W = fL.shape
    data = ConcatVolumePaddedFunction.apply(fL, fR, int(D), _lib.VOLUME_MODES[mode])
    return PaddedVolume(data, B, 2 * C, int(D), H, W)


def concat_volume(fL, fR, D, mode="psm", padded_bf16=False, out=None):
    if fL.shape != fR.shape or fL.dim() != 4:
        raise _lib.DsmError("concat_volume expects two NCHW tensors of equal shape")
    m = _lib.VOLUME_MODES[mode]
    D = int(D)
    if not padded_bf16:
        return ConcatVolumeFunction.apply(fL, fR, D, m)
    _lib.require_cuda(fL, fR)
    fL = fL.contiguous().float(); fR = fR.contiguous().float()
    B, C, H, W = fL.shape
    vol = out if out is not None else PaddedVolume.empty(B, 2 * C, D, H, W, fL.device, zero_rim=False)
    assert vol.shape5 == (B, 2 * C, D, H, W)
    _lib.check(_lib.lib().dsm_concat_volume_fwd(fL.data_ptr(), fR.data_ptr(), vol.data.data_ptr(), B, C, D, H, W, m,
                                                _lib.DSM_BF16, _lib.DSM_NDHWC_PADDED, _lib.stream_ptr(fL.device)),
               "dsm_concat_volume_fwd")
    return vol
