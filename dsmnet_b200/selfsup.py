"""Self-supervised training step of the reference (callers on both sides of op 5, the imwrap warp):
the `depthmono[-mask]` pyramid loss (losses/loss.py:196-236 `loss_depthmono`, :388-404 `weight_common`,
:424-466 `losses_pyramid1`, losses/SSIM.py:24-42) and the step of stereo_selfsupervised.py:48-100
(mirrored pair, two forwards, loss, backward), plus the multi-GPU form of it: one process per GPU,
batch sharded by rank, NCCL gradient all-reduce through DistributedDataParallel.

Hot-path ops inside: 28 `imwrap_BCHW` warps per step (7 levels x 4) — ONE batched launch forward and one backward
(`imwrap.imwrap_batched`) — and DispNetC's 1-D correlation, on the sm_100a kernels with their backward kernels; the
SSIM map (five 11x11 Gaussian convolutions + ~15 elementwise kernels in the reference) is one fused kernel forward and
two backward (`SsimFunction`, csrc/ssim.cu); the mirrored pair is a device flip.  Smoothness / masks / weights are stock
elementwise ops.  `warp_fn` is injectable so that the host logic of the loss is testable on CPU with the oracle's
warp (tests/test_selfsup_cpu.py); the default is the CUDA op and has no fallback.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import torch
import torch.nn.functional as F

_WINDOWS = {}


def _gauss_window(size: int, channel: int, like: torch.Tensor) -> torch.Tensor:
    """create_window (SSIM.py:10-14) in the form `_ssim` uses it (:25): [1, C, k, k] / C."""
    key = (size, channel, like.device, like.dtype)
    w = _WINDOWS.get(key)
    if w is None:
        g = torch.tensor([math.exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])
        g = g / g.sum()
        w2 = g.unsqueeze(1).mm(g.unsqueeze(0)).float()
        w = (w2.expand(channel, 1, size, size).contiguous().transpose(0, 1) / channel).to(like.device, like.dtype)
        _WINDOWS[key] = w
    return w


class SsimFunction(torch.autograd.Function):
    """SSIM.py:24-42 as one fused kernel (dsm_ssim_fwd); the gradient flows to b (the warped image) through two more
    (dsm_ssim_bwd).  `a` is the real image: asking for its gradient raises (the reference's loss never does)."""

    @staticmethod
    def forward(ctx, a, b):
        from . import _lib
        _lib.require_cuda(a, b)
        a = a.contiguous().float(); b = b.contiguous().float()
        B, C, H, W = a.shape
        s = torch.empty(B, 1, H, W, device=a.device, dtype=torch.float32)
        _lib.check(_lib.lib().dsm_ssim_fwd(a.data_ptr(), b.data_ptr(), s.data_ptr(), B, C, H, W, _lib.stream_ptr(a.device)), "dsm_ssim_fwd")
        ctx.save_for_backward(a, b)
        return s

    @staticmethod
    def backward(ctx, gs):
        from . import _lib
        a, b = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise _lib.DsmError("SsimFunction: gradient w.r.t. the first (real) image is not implemented")
        B, C, H, W = a.shape
        gs = gs.contiguous().float()
        gb = torch.empty_like(b)
        L = _lib.lib()
        ws = torch.empty(L.dsm_ssim_bwd_workspace_bytes(B, H, W) // 4, device=a.device, dtype=torch.float32)
        _lib.check(L.dsm_ssim_bwd(a.data_ptr(), b.data_ptr(), gs.data_ptr(), gb.data_ptr(), B, C, H, W, ws.data_ptr(), ws.numel() * 4,
                                  _lib.stream_ptr(a.device)), "dsm_ssim_bwd")
        return None, gb


def ssim_map(a: torch.Tensor, b: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """SSIM.py:24-42 (`_ssim`): channel-averaged 11x11 Gaussian SSIM map, [B,1,H,W].  CUDA tensors take the fused kernels
    (`SsimFunction`); CPU tensors (the host-logic tests with the oracle's warp) the stock restatement below."""
    if a.is_cuda and window_size == 11 and not a.requires_grad:
        return SsimFunction.apply(a, b)
    w = _gauss_window(window_size, a.shape[1], a)
    p = window_size // 2
    mu1, mu2 = F.conv2d(a, w, padding=p), F.conv2d(b, w, padding=p)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(a * a, w, padding=p) - mu1_sq
    s2 = F.conv2d(b * b, w, padding=p) - mu2_sq
    s12 = F.conv2d(a * b, w, padding=p) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def _dx(t):
    return F.pad(t[:, :, :, 1:] - t[:, :, :, :-1], [0, 1, 0, 0])


def _dy(t):
    return F.pad(t[:, :, 1:] - t[:, :, :-1], [0, 0, 0, 1])


def smoothness_ds1(im: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """C_ds1 (loss.py:71-83): edge-aware first-order disparity smoothness."""
    wx = torch.exp(-torch.sum(torch.abs(_dx(im)), dim=1, keepdim=True))
    wy = torch.exp(-torch.sum(torch.abs(_dy(im)), dim=1, keepdim=True))
    return torch.abs(_dx(disp)) * wx + torch.abs(_dy(disp)) * wy


def weight_common(disp: torch.Tensor, disp_wrap: torch.Tensor, factor: float = 1.0) -> torch.Tensor:
    """loss.py:388-404: left-right consistency weight, 1 below 1 px, linear to 0.01 at 3 px."""
    d = torch.abs(disp - disp_wrap).detach() / factor
    w = torch.full_like(d, 0.01)
    w = torch.where(d < 3, 1.0 - (d - 1) * (0.99 / 2), w)
    return torch.where(d < 1, torch.ones_like(d), w)


def loss_depthmono(im, im_wrap, disp, disp_wrap, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """loss.py:196-236: appearance (0.85 SSIM + 0.15 L1) + w * smoothness + w * left-right, with
    w = max(0, mean SSIM over the warp's valid mask - 0.75)/2 + 0.001 (:27-28, :202-205)."""
    s = ssim_map(im, im_wrap)
    mask_ap = (im_wrap[:, :1] != 0).detach()
    if int(mask_ap.sum()) < 1024:
        mask_ap = torch.ones_like(mask_ap)
    similarity = float(s.detach()[mask_ap].mean())
    w = max(0.0, similarity - 0.75) / 2 + 0.001
    c_ap = (0.85 * 0.5) * (1 - s) + 0.15 * torch.abs(im - im_wrap)
    c_lr = torch.abs(disp - disp_wrap)
    if weight is not None:
        zero = (disp_wrap == 0)
        c_ap = c_ap * torch.where((zero & mask_ap).detach(), torch.ones_like(weight), weight)
        c_lr = c_lr * torch.where(zero, torch.zeros_like(weight), weight)
    return c_ap.mean() * 1.0 + smoothness_ds1(im, disp).mean() * w + c_lr.mean() * w


def _default_warp():
    from .imwrap import imwrap_BCHW
    return imwrap_BCHW


def losses_pyramid1(imR_src, imL, dispLs: Sequence[torch.Tensor], scales: Sequence[int], LeftTop,
                    imR1_src, imL1, dispL1s: Sequence[torch.Tensor], LeftTop1, weight_levels: Sequence[float],
                    flag_mask: bool = True, warp_fn: Optional[Callable] = None, align_corners: bool = True) -> torch.Tensor:
    """loss.py:424-466.  Levels above 2 are upsampled to level 2; four warps per level: the two photometric warps of the
    (cropped) left images from the uncropped right sources, and the two left-right disparity warps (fliplr)."""
    batched = warp_fn is None          # CUDA default: the 4 x len(scales) warps of this call as ONE launch (fwd) + one (bwd)
    warp = warp_fn or _default_warp()
    maxlevel = min(2, max(scales))
    h, w = dispLs[list(scales).index(maxlevel)].shape[-2:]
    imLs, imL1s = [imL], [imL1]
    for _ in range(maxlevel):
        imLs.append(imLs[-1][:, :, ::2, ::2]); imL1s.append(imL1s[-1][:, :, ::2, ::2])
    loss = 0
    levels = []
    for i, level in enumerate(scales):
        wl = weight_levels[level]
        if wl <= 0:
            continue
        if level > maxlevel:
            sf = 2 ** maxlevel
            up = lambda t: F.interpolate(t, scale_factor=2 ** (level - maxlevel), mode="bilinear", align_corners=align_corners)[:, :, :h, :w]
            dL, dL1 = up(dispLs[i]), up(dispL1s[i])
        else:
            sf = 2 ** level
            dL, dL1 = dispLs[i], dispL1s[i]
        levels.append((level, wl, sf, dL, dL1))
    warped = None
    if batched and levels:
        from .imwrap import imwrap_batched
        jobs = []
        for level, wl, sf, dL, dL1 in levels:                       # the reference's call order (its RNG draws follow it)
            jobs += [dict(im_src=imR_src, disp=dL, fliplr=False, LeftTop=list(LeftTop), scale_factor=sf),
                     dict(im_src=imR1_src, disp=dL1, fliplr=False, LeftTop=list(LeftTop1), scale_factor=sf),
                     dict(im_src=dL1, disp=dL, fliplr=True, LeftTop=[0, 0], scale_factor=1),
                     dict(im_src=dL, disp=dL1, fliplr=True, LeftTop=[0, 0], scale_factor=1)]
        warped = imwrap_batched(jobs)
    for n, (level, wl, sf, dL, dL1) in enumerate(levels):
        if warped is not None:
            imL_w, imL1_w, dL_w, dL1_w = warped[4 * n:4 * n + 4]
        else:
            imL_w = warp(imR_src, dL, fliplr=False, LeftTop=list(LeftTop), scale_factor=sf)
            imL1_w = warp(imR1_src, dL1, fliplr=False, LeftTop=list(LeftTop1), scale_factor=sf)
            dL_w = warp(dL1, dL, fliplr=True, LeftTop=[0, 0], scale_factor=1)
            dL1_w = warp(dL, dL1, fliplr=True, LeftTop=[0, 0], scale_factor=1)
        wc = weight_common(dL, dL_w, sf) if flag_mask else None
        wc1 = weight_common(dL1, dL1_w, sf) if flag_mask else None
        k = min(level, maxlevel)
        loss = loss + (loss_depthmono(imLs[k], imL_w, dL, dL_w, wc) + loss_depthmono(imL1s[k], imL1_w, dL1, dL1_w, wc1)) * wl
    return loss


def flip_lr(t: torch.Tensor) -> torch.Tensor:
    """flip_lr_tensor (stereo_selfsupervised.py:44-46) without the numpy round trip."""
    return torch.flip(t, dims=[3])


def selfsup_loss_for_batch(model, batch: torch.Tensor, nedge: int, weight_levels: Sequence[float], flag_mask: bool = True,
                           warp_fn: Optional[Callable] = None) -> torch.Tensor:
    """Forward part of stereo_selfsupervised.py:60-95 for a [B, 6, h, w] batch (left RGB | right RGB), without the colour
    augmentation (a data-side transform): crop by `nedge`, mirrored pair, two model forwards, pyramid loss."""
    b1 = flip_lr(batch)
    crop = lambda t: t[:, :, nedge:t.shape[2] - nedge, nedge:t.shape[3] - nedge]
    imL, imR = crop(batch[:, :3]).contiguous(), crop(batch[:, 3:6]).contiguous()
    imL1, imR1 = crop(b1[:, 3:6]).contiguous(), crop(b1[:, :3]).contiguous()
    scales, dispLs = model(imL, imR)
    _, dispL1s = model(imL1, imR1)
    return losses_pyramid1(batch[:, 3:6], imL, dispLs, scales, (nedge, nedge), b1[:, :3], imL1, dispL1s, (nedge, nedge),
                           weight_levels, flag_mask, warp_fn)


def train_step(model, optimizer, batch: torch.Tensor, nedge: int, weight_levels: Sequence[float],
               warp_fn: Optional[Callable] = None) -> torch.Tensor:
    """One optimisation step (stereo_selfsupervised.py:60-100).  With `model` wrapped in DistributedDataParallel the
    backward all-reduces (averages) the gradients over NCCL (NVLink/NVSwitch), bucketed and overlapped with the backward
    pass; each rank passes ITS shard of the global batch (dsmnet_b200.shard.shard_range)."""
    optimizer.zero_grad(set_to_none=True)
    loss = selfsup_loss_for_batch(model, batch, nedge, weight_levels, True, warp_fn)
    loss.backward()
    optimizer.step()
    return loss.detach()
