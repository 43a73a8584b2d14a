"""Drop-in for the reference's ``Corr1d`` (models/util_conv.py:56-86) on the sm_100a kernel.

Same constructor and call signature: ``Corr1d(kernel_size=1, stride=1, D=1, simfun=None)(fL, fR)``
with NCHW fp32 inputs and a ``(B, D, H, W)`` output.  A non-default ``simfun`` raises (the
reference's call sites, dispnetcorr.py:27 and iresnet.py:34,69, all pass None) — there is no
slow path.  The k>1 average pool of :82-85 stays the stock pooling op the reference uses.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class Corr1dFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fL, fR, D, stride):
        _lib.require_cuda(fL, fR)
        if fL.shape != fR.shape or fL.dim() != 4:
            raise _lib.DsmError("Corr1d expects two NCHW tensors of equal shape")
        fL = fL.contiguous().float(); fR = fR.contiguous().float()
        B, C, H, W = fL.shape
        out = torch.empty(B, D, H, W, device=fL.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_corr1d_fwd(fL.data_ptr(), fR.data_ptr(), out.data_ptr(), B, C, H, W, D, stride,
                                             _lib.stream_ptr(fL.device)), "dsm_corr1d_fwd")
        ctx.save_for_backward(fL, fR)
        ctx.D, ctx.stride = D, stride
        return out

    @staticmethod
    def backward(ctx, g):
        fL, fR = ctx.saved_tensors
        g = g.contiguous().float()
        B, C, H, W = fL.shape
        gL = torch.empty_like(fL); gR = torch.empty_like(fR)
        _lib.check(_lib.lib().dsm_corr1d_bwd(g.data_ptr(), fL.data_ptr(), fR.data_ptr(), gL.data_ptr(), gR.data_ptr(),
                                             B, C, H, W, ctx.D, ctx.stride, _lib.stream_ptr(fL.device)), "dsm_corr1d_bwd")
        return gL, gR, None, None


def corr1d(fL, fR, D, stride=1):
    return Corr1dFunction.apply(fL, fR, int(D), int(stride))


class Corr1d(nn.Module):
    def __init__(self, kernel_size=1, stride=1, D=1, simfun=None):
        super().__init__()
        if simfun is not None:
            raise _lib.DsmError("Corr1d: only the default dot-product simfun (util_conv.py:68-69) is implemented")
        self.kernel_size, self.stride, self.D = kernel_size, stride, D

    def forward(self, fL, fR):
        corrmap = corr1d(fL, fR, self.D, self.stride)
        if self.kernel_size > 1:
            assert self.kernel_size % 2 == 1
            corrmap = F.avg_pool2d(corrmap, self.kernel_size, stride=1, padding=self.kernel_size // 2)
        return corrmap
