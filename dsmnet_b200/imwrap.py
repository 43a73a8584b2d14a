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


# --------------------------------------------------------------------------------------------------------------------
# All the warps of one training step in one launch (forward) and one launch (backward): losses/loss.py:449-452 issues four
# imwrap_BCHW calls per pyramid level, 28 per step, none of which reads another one's output.
# --------------------------------------------------------------------------------------------------------------------

class BatchedWarpFunction(torch.autograd.Function):
    """outs = [imwrap_BCHW(src_i, disp_i, fliplr_i, LeftTop_i, scale_i) for i]; `meta[i]` = (fliplr, LeftTop, scale_factor, delt).
    tensors = (src_0, disp_0, src_1, disp_1, ...).  One dsm_warp_fwd_batched launch; backward: one memset + one launch."""

    @staticmethod
    def _jobs(meta, tensors, outs=None, gouts=None, gsrcs=None, gdisps=None):
        n = len(meta)
        arr = (_lib.DsmWarpJob * n)()
        for i, (fliplr, lefttop, sf, delt) in enumerate(meta):
            src, disp = tensors[2 * i], tensors[2 * i + 1]
            B, C, H0, W0 = src.shape
            _, _, H, W = disp.shape
            row, col = grid_vectors_device(H0, W0, H, W, lefttop, sf, src.device)
            j = arr[i]
            j.src, j.disp, j.row, j.col = src.data_ptr(), disp.data_ptr(), row.data_ptr(), col.data_ptr()
            j.out = outs[i].data_ptr() if outs is not None else 0
            j.gout = gouts[i].data_ptr() if gouts is not None else 0
            j.gsrc = gsrcs[i].data_ptr() if (gsrcs is not None and gsrcs[i] is not None) else 0
            j.gdisp = gdisps[i].data_ptr() if gdisps is not None else 0
            j.delt, j.fliplr = float(delt), int(bool(fliplr))
            j.B, j.C, j.H0, j.W0, j.H, j.W = B, C, H0, W0, H, W
        return arr

    @staticmethod
    def forward(ctx, meta, *tensors):
        _lib.require_cuda(*tensors)
        tensors = tuple(t.contiguous().float() for t in tensors)
        n = len(meta)
        if n < 1 or n > _lib.WARP_MAX_JOBS or len(tensors) != 2 * n:
            raise _lib.DsmError("BatchedWarpFunction: 1..%d jobs, two tensors each" % _lib.WARP_MAX_JOBS)
        outs = []
        for i in range(n):
            src, disp = tensors[2 * i], tensors[2 * i + 1]
            assert disp.shape[1] == 1 and min(src.shape[2], src.shape[3], disp.shape[2], disp.shape[3]) > 1
            outs.append(torch.empty(src.shape[0], src.shape[1], disp.shape[2], disp.shape[3], device=src.device, dtype=torch.float32))
        jobs = BatchedWarpFunction._jobs(meta, tensors, outs=outs)
        _lib.check(_lib.lib().dsm_warp_fwd_batched(jobs, n, _lib.stream_ptr(tensors[0].device)), "dsm_warp_fwd_batched")
        ctx.save_for_backward(*tensors)
        ctx.meta = meta
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tensors = ctx.saved_tensors
        meta = ctx.meta
        n = len(meta)
        dev = tensors[0].device
        g = [(torch.zeros(tensors[2 * i].shape[0], tensors[2 * i].shape[1], *tensors[2 * i + 1].shape[2:], device=dev) if gouts[i] is None
              else gouts[i].contiguous().float()) for i in range(n)]
        # one zeroed allocation for every source gradient that is wanted (images need none)
        want = [ctx.needs_input_grad[1 + 2 * i] for i in range(n)]
        sizes = [tensors[2 * i].numel() if want[i] else 0 for i in range(n)]
        pool = torch.zeros(max(1, sum(sizes)), device=dev, dtype=torch.float32)
        gsrcs, off = [], 0
        for i in range(n):
            gsrcs.append(pool[off:off + sizes[i]].view_as(tensors[2 * i]) if want[i] else None)
            off += sizes[i]
        gdisps = [torch.empty_like(tensors[2 * i + 1]) for i in range(n)]
        jobs = BatchedWarpFunction._jobs(meta, tensors, gouts=g, gsrcs=gsrcs, gdisps=gdisps)
        _lib.check(_lib.lib().dsm_warp_bwd_batched(jobs, n, _lib.stream_ptr(dev)), "dsm_warp_bwd_batched")
        grads = [None]
        for i in range(n):
            grads += [gsrcs[i], gdisps[i] if ctx.needs_input_grad[2 + 2 * i] else None]
        return tuple(grads)


def imwrap_batched(jobs):
    """jobs: list of dicts(im_src, disp, fliplr=False, LeftTop=[0, 0], scale_factor=1, delt=None) -> list of warped tensors.
    Draws one torch.rand(1) per job, in order, exactly like consecutive imwrap_BCHW calls (imwrap.py:70)."""
    meta, tensors = [], []
    for j in jobs:
        delt = j.get("delt")
        if delt is None:
            delt = float(1e-4 * (torch.rand(1)[0] + 0.1))
        meta.append((bool(j.get("fliplr", False)), tuple(j.get("LeftTop", (0, 0))), j.get("scale_factor", 1), delt))
        tensors += [j["im_src"], j["disp"]]
    return list(BatchedWarpFunction.apply(meta, *tensors))
