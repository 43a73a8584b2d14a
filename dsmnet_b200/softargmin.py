"""Soft-argmin disparity regression on the sm_100a kernels.

``disparityregression(maxdisp)(x)`` mirrors models/psmnet/submodule.py:56-63 (input: softmaxed
probabilities [B,D,H,W]; output [B,H,W] without channel dim).  ``softargmin(cost, sign)`` fuses
the softmax that precedes it (stackhourglass.py:155-157; GC-Net gcnet.py:104-109 with sign=-1),
``upsample_softargmin(cost_lr, (D,H,W))`` additionally fuses the trilinear upsample of
stackhourglass.py:152-153,163 (inference only).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class SoftArgminFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, sign):
        _lib.require_cuda(cost)
        cost = cost.contiguous().float()
        B, D, H, W = cost.shape
        disp = torch.empty(B, H, W, device=cost.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_softargmin_fwd(cost.data_ptr(), disp.data_ptr(), B, D, H, W, float(sign),
                                                 _lib.stream_ptr(cost.device)), "dsm_softargmin_fwd")
        ctx.save_for_backward(cost, disp)
        ctx.sign = float(sign)
        return disp

    @staticmethod
    def backward(ctx, g):
        cost, disp = ctx.saved_tensors
        B, D, H, W = cost.shape
        g = g.contiguous().float()
        gc = torch.empty_like(cost)
        _lib.check(_lib.lib().dsm_softargmin_bwd(cost.data_ptr(), disp.data_ptr(), g.data_ptr(), gc.data_ptr(), B, D, H, W,
                                                 ctx.sign, _lib.stream_ptr(cost.device)), "dsm_softargmin_bwd")
        return gc, None


class DisparityRegressionFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob):
        _lib.require_cuda(prob)
        prob = prob.contiguous().float()
        B, D, H, W = prob.shape
        disp = torch.empty(B, H, W, device=prob.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_disparity_regression_fwd(prob.data_ptr(), disp.data_ptr(), B, D, H, W,
                                                           _lib.stream_ptr(prob.device)), "dsm_disparity_regression_fwd")
        ctx.shape = (B, D, H, W)
        return disp

    @staticmethod
    def backward(ctx, g):
        B, D, H, W = ctx.shape
        g = g.contiguous().float()
        gp = torch.empty(B, D, H, W, device=g.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_disparity_regression_bwd(g.data_ptr(), gp.data_ptr(), B, D, H, W,
                                                           _lib.stream_ptr(g.device)), "dsm_disparity_regression_bwd")
        return gp


def softargmin(cost, sign=1.0):
    """[B,D,H,W] cost -> [B,H,W] expected disparity under softmax_d(sign*cost)."""
    return SoftArgminFunction.apply(cost, sign)


class UpsampleSoftargminFunction(torch.autograd.Function):
    """Differentiable fused head (stackhourglass.py:152-166): trilinear upsample -> softmax over D -> regression,
    forward and backward without the upsampled [B, D, H, W] volume (dsm_upsample_softargmin_fwd_lse / _bwd)."""

    @staticmethod
    def forward(ctx, cost_lr, size, align_corners):
        _lib.require_cuda(cost_lr)
        cost_lr = cost_lr.contiguous().float()
        B, Dl, Hl, Wl = cost_lr.shape
        D, H, W = (int(s) for s in size)
        disp = torch.empty(B, H, W, device=cost_lr.device, dtype=torch.float32)
        lse2 = torch.empty(B, H, W, device=cost_lr.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_upsample_softargmin_fwd_lse(cost_lr.data_ptr(), disp.data_ptr(), lse2.data_ptr(), B, Dl, Hl, Wl,
                                                              D, H, W, 1 if align_corners else 0, _lib.stream_ptr(cost_lr.device)),
                   "dsm_upsample_softargmin_fwd_lse")
        ctx.save_for_backward(cost_lr, disp, lse2)
        ctx.size, ctx.align_corners = (D, H, W), bool(align_corners)
        return disp

    @staticmethod
    def backward(ctx, gdisp):
        cost_lr, disp, lse2 = ctx.saved_tensors
        B, Dl, Hl, Wl = cost_lr.shape
        D, H, W = ctx.size
        gdisp = gdisp.contiguous().float()
        gcost = torch.empty_like(cost_lr)
        _lib.check(_lib.lib().dsm_upsample_softargmin_bwd(cost_lr.data_ptr(), disp.data_ptr(), lse2.data_ptr(), gdisp.data_ptr(),
                                                          gcost.data_ptr(), B, Dl, Hl, Wl, D, H, W, 1 if ctx.align_corners else 0,
                                                          _lib.stream_ptr(cost_lr.device)), "dsm_upsample_softargmin_bwd")
        return gcost, None, None


def upsample_softargmin(cost_lr, size, align_corners=True, out=None):
    """cost_lr [B,Dl,Hl,Wl] (or [B,1,Dl,Hl,Wl]) -> disparity [B,H,W] at ``size=(D,H,W)``.
    Differentiable (fused backward kernel) when cost_lr requires grad; `out` is honoured on the no-grad route only."""
    _lib.require_cuda(cost_lr)
    if cost_lr.dim() == 5:
        cost_lr = cost_lr.squeeze(1)
    if torch.is_grad_enabled() and cost_lr.requires_grad:
        return UpsampleSoftargminFunction.apply(cost_lr, tuple(int(s) for s in size), bool(align_corners))
    cost_lr = cost_lr.contiguous().float()
    B, Dl, Hl, Wl = cost_lr.shape
    D, H, W = (int(s) for s in size)
    disp = out if out is not None else torch.empty(B, H, W, device=cost_lr.device, dtype=torch.float32)
    _lib.check(_lib.lib().dsm_upsample_softargmin_fwd(cost_lr.data_ptr(), disp.data_ptr(), B, Dl, Hl, Wl, D, H, W,
                                                      1 if align_corners else 0, _lib.stream_ptr(cost_lr.device)),
               "dsm_upsample_softargmin_fwd")
    return disp


class disparityregression(nn.Module):
    """Same name/ctor/call as the reference module (submodule.py:56-63)."""

    def __init__(self, maxdisp):
        super().__init__()
        self.maxdisp = maxdisp

    def forward(self, x):
        return DisparityRegressionFunction.apply(x)
