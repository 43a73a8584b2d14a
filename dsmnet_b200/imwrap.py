"""Drop-in for the reference's ``imwrap_BCHW`` (utils/imwrap.py:37-72) on the sm_100a kernel.

Same signature and semantics: the sampling-grid vectors are built on the host with
``torch.linspace`` exactly as imwrap.py:50-58 does, one ``torch.rand(1)`` is drawn from the
global CPU generator per call (imwrap.py:70) so the RNG stream stays identical, and the device
part (disparity shift, ``+delt``, bilinear ``grid_sample`` with align_corners=True semantics) is
one fused kernel with gradients to both ``im_src`` and ``disp``.
"""
from __future__ import annotations

import torch

from . import _lib


class WarpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im_src, disp, row, col, delt, fliplr):
        _lib.require_cuda(im_src, disp, row, col)
        im_src = im_src.contiguous().float(); disp = disp.contiguous().float()
        B, C, H0, W0 = im_src.shape
        _, _, H, W = disp.shape
        out = torch.empty(B, C, H, W, device=im_src.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_warp_fwd(im_src.data_ptr(), disp.data_ptr(), row.data_ptr(), col.data_ptr(),
                                           float(delt), int(bool(fliplr)), out.data_ptr(), B, C, H0, W0, H, W,
                                           _lib.stream_ptr(im_src.device)), "dsm_warp_fwd")
        ctx.save_for_backward(im_src, disp, row, col)
        ctx.delt, ctx.fliplr = float(delt), int(bool(fliplr))
        return out

    @staticmethod
    def backward(ctx, g):
        im_src, disp, row, col = ctx.saved_tensors
        B, C, H0, W0 = im_src.shape
        _, _, H, W = disp.shape
        g = g.contiguous().float()
        gsrc = torch.empty_like(im_src)
        gdisp = torch.empty_like(disp)
        _lib.check(_lib.lib().dsm_warp_bwd(g.data_ptr(), im_src.data_ptr(), disp.data_ptr(), row.data_ptr(), col.data_ptr(),
                                           ctx.delt, ctx.fliplr, gsrc.data_ptr(), gdisp.data_ptr(), B, C, H0, W0, H, W,
                                           _lib.stream_ptr(im_src.device)), "dsm_warp_bwd")
        return gsrc, gdisp, None, None, None, None


def grid_vectors(h0, w0, h, w, LeftTop=(0, 0), scale_factor=1):
    """imwrap.py:50-58, verbatim arithmetic (python floats, then torch.linspace in fp32)."""
    x, y = LeftTop
    x = x * 2.0 / (w0 - 1) - 1
    y = y * 2.0 / (h0 - 1) - 1
    x1 = x + (w - 1) * scale_factor * 2.0 / (w0 - 1)
    y1 = y + (h - 1) * scale_factor * 2.0 / (h0 - 1)
    return torch.linspace(x, x1, w), torch.linspace(y, y1, h)


_grid_cache = {}


def grid_vectors_device(h0, w0, h, w, LeftTop, scale_factor, device):
    """The two linspace vectors on `device`, built on the host once per geometry and kept (a per-call host->device copy of
    pageable memory would make the op uncapturable in a CUDA graph and costs more than the kernel for small levels)."""
    key = (h0, w0, h, w, float(LeftTop[0]), float(LeftTop[1]), float(scale_factor), str(device))
    rc = _grid_cache.get(key)
    if rc is None:
        if len(_grid_cache) > 256:
            _grid_cache.clear()
        row, col = grid_vectors(h0, w0, h, w, LeftTop, scale_factor)
        rc = _grid_cache[key] = (row.to(device), col.to(device))
    return rc


def imwrap_BCHW(im_src, disp, fliplr=False, LeftTop=[0, 0], scale_factor=1, delt=None):
    """``delt=None`` draws it like the reference (1e-4*(U[0,1)+0.1) from the global CPU RNG)."""
    bn, _, h0, w0 = im_src.shape
    bn, c, h, w = disp.shape
    assert c == 1 and min(h, w, h0, w0) > 1
    _lib.require_cuda(im_src, disp)
    row, col = grid_vectors_device(h0, w0, h, w, LeftTop, scale_factor, im_src.device)
    if delt is None:
        delt = float(1e-4 * (torch.rand(1)[0] + 0.1))
    return WarpFunction.apply(im_src, disp, row, col, delt, fliplr)


def imwrap_pyramid(im_src, disps_pyramid, fliplr=False, LeftTop=[0, 0]):
    """utils/imwrap.py:26-35."""
    assert type(disps_pyramid) is list
    ims_wrap = []
    scale_factor = 1
    for d in disps_pyramid:
        ims_wrap.append(imwrap_BCHW(im_src, d, fliplr, LeftTop, scale_factor))
        scale_factor = scale_factor * 2
    return ims_wrap
