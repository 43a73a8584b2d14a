"""iResNet on the sm_100a hot-path kernels — drop-in for the reference's ``iresnet``
(models/iresnet.py:17-200; BASELINE.json config 4).

Hot-path ops inside: ``Corr1d(D=81)`` on the 1/4-resolution features (iresnet.py:34,107), the
feature-constancy warp ``imwrap_BCHW(deconv1R2R, -r_pr0)`` at full resolution (:169) and
``Corr1d(kernel_size=3, stride=2, D=41)`` on the 1/2-resolution features (:69,175).  Everything else
is a stock 2-D conv / deconv + bias + ReLU exactly as in the reference and stays cuDNN.  The module
is generated from layer tables; parameter names are the reference's, so its ``state_dict`` loads
unchanged.  ``corr_fn`` / ``warp_fn`` are injectable so that the graph can be checked on CPU against
the reference with the oracle's ops (tests/test_models_cpu.py); the defaults are the CUDA ops.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .dispnetcorr import crop_cat

# name, kind ("c" conv+ReLU, "d" deconv+ReLU, "p" bare prediction conv), Cin, Cout, kernel, stride
LAYERS = [
    ("conv1", "c", 3, 64, 7, 2), ("conv2", "c", 64, 128, 5, 2),
    ("deconv1_s", "d", 64, 32, 4, 2), ("deconv2_s", "d", 128, 32, 8, 4), ("conv_de1_de2", "c", 64, 32, 1, 1),
    ("redir", "c", 128, 64, 1, 1), ("conv3", "c", 81 + 64, 256, 3, 2), ("conv3_1", "c", 256, 256, 3, 1),
    ("conv4", "c", 256, 512, 3, 2), ("conv4_1", "c", 512, 512, 3, 1), ("conv5", "c", 512, 512, 3, 2), ("conv5_1", "c", 512, 512, 3, 1),
    ("conv6", "c", 512, 1024, 3, 2), ("conv6_1", "c", 1024, 1024, 3, 1), ("pr6", "p", 1024, 1, 3, 1),
    ("deconv5", "d", 1024, 512, 4, 2), ("iconv5", "c", 1025, 512, 3, 1), ("pr5", "p", 512, 1, 3, 1),
    ("deconv4", "d", 512, 256, 4, 2), ("iconv4", "c", 769, 256, 3, 1), ("pr4", "p", 256, 1, 3, 1),
    ("deconv3", "d", 256, 128, 4, 2), ("iconv3", "c", 385, 128, 3, 1), ("pr3", "p", 128, 1, 3, 1),
    ("deconv2", "d", 128, 64, 4, 2), ("iconv2", "c", 193, 64, 3, 1), ("pr2", "p", 64, 1, 3, 1),
    ("deconv1", "d", 64, 32, 4, 2), ("iconv1", "c", 97, 32, 3, 1), ("pr1", "p", 32, 1, 3, 1),
    ("deconv0", "d", 32, 32, 4, 2), ("iconv0", "c", 65, 32, 3, 1), ("pr0", "p", 32, 1, 3, 1),
    ("r_conv0", "c", 65, 32, 3, 1), ("r_conv1", "c", 32, 64, 3, 2), ("c_conv1", "c", 64, 64, 3, 1),
    ("r_conv1_1", "c", 105, 64, 3, 1), ("r_conv2", "c", 64, 128, 3, 2), ("r_conv2_1", "c", 128, 128, 3, 1), ("r_res2", "p", 128, 1, 3, 1),
    ("r_deconv1", "d", 128, 64, 4, 2), ("r_iconv1", "c", 129, 64, 3, 1), ("r_res1", "p", 64, 1, 3, 1),
    ("r_deconv0", "d", 64, 32, 4, 2), ("r_iconv0", "c", 65, 32, 3, 1), ("r_res0", "p", 32, 1, 3, 1),
]


def _make(kind, cin, cout, k, s):
    if kind == "p":
        return nn.Conv2d(cin, cout, k, s, padding=(k - 1) // 2)
    if kind == "c":
        return nn.Sequential(nn.Conv2d(cin, cout, k, s, padding=(k - 1) // 2, bias=True), nn.ReLU(inplace=True))
    p = (k - 1) // 2                                            # deconv2d_bn (util_conv.py:133-148)
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, k, s, padding=p, output_padding=s - (k - 2 * p), bias=True), nn.ReLU(inplace=True))


class iresnet(nn.Module):
    def __init__(self, maxdisparity=192, align_corners=True, corr_fn=None, warp_fn=None):
        super().__init__()
        self.name = "iresnet"
        self.D = maxdisparity
        self.delt = 1e-6
        self.count_levels = 7
        self.align_corners = align_corners
        for name, kind, cin, cout, k, s in LAYERS:
            setattr(self, name, _make(kind, cin, cout, k, s))
        for m in self.modules():                                   # net_init (util_conv.py:32-53)
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
        for name, kind, *_ in LAYERS:                               # iresnet.py:84-85 (pr0 is not in that list)
            if kind == "p" and name != "pr0":
                getattr(self, name).weight.data.mul_(0.1)
        self._corr_fn, self._warp_fn = corr_fn, warp_fn

    # the three hot-path call sites ------------------------------------------------------------------------------
    def _corr(self, a, b, kernel_size, stride, D):
        if self._corr_fn is not None:
            return self._corr_fn(a, b, D, stride, kernel_size)
        from .corr1d import corr1d
        c = corr1d(a, b, D, stride)
        return F.avg_pool2d(c, kernel_size, stride=1, padding=kernel_size // 2) if kernel_size > 1 else c

    def _warp(self, src, disp):
        if self._warp_fn is not None:
            return self._warp_fn(src, disp)
        from .imwrap import imwrap_BCHW
        return imwrap_BCHW(src, disp)

    def _up(self, x):
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=self.align_corners)

    def forward(self, imL, imR, mode="train", iter=1):
        assert imL.shape == imR.shape
        H, W = imL.shape[-2:]
        maxD = max(self.D, W)
        c1L, c1R = self.conv1(imL), self.conv1(imR)
        c2L, c2R = self.conv2(c1L), self.conv2(c1R)
        # multi-scale shared features at full resolution (:92-101)
        fuse = lambda c1, c2: self.conv_de1_de2(crop_cat(self.deconv1_s(c1)[:, :, :H, :W], self.deconv2_s(c2)))
        featL, featR = fuse(c1L, c2L), fuse(c1R, c2R)
        # initial disparity: DispNetC-like encoder/decoder on the D=81 correlation (:104-166)
        x = torch.cat([self._corr(c2L, c2R, 1, 1, 81), self.redir(c2L)], dim=1)
        skips = {0: featL, 1: c1L, 2: c2L}
        for lvl in (3, 4, 5, 6):
            x = getattr(self, "conv%d_1" % lvl)(getattr(self, "conv%d" % lvl)(x))
            skips[lvl] = x
        out, scales = [self.pr6(x)], [6]
        for lvl in (5, 4, 3, 2, 1, 0):
            x = getattr(self, "iconv%d" % lvl)(crop_cat(getattr(self, "deconv%d" % lvl)(x), self._up(out[0]), skips[lvl]))
            out.insert(0, getattr(self, "pr%d" % lvl)(x)); scales.insert(0, lvl)
        r_pr = {0: out[0], 1: out[1], 2: out[2]}
        # iterative refinement: feature constancy (warp) + a second, half-resolution correlation (:168-195)
        for _ in range(iter):
            recon = torch.abs(featL - self._warp(featR, -r_pr[0]))
            r0 = self.r_conv0(crop_cat(recon, r_pr[0], featL))
            r1 = self.r_conv1(r0)
            r1 = self.r_conv1_1(crop_cat(r1, self._corr(self.c_conv1(c1L), self.c_conv1(c1R), 3, 2, 41)))
            r2 = self.r_conv2_1(self.r_conv2(r1))
            res2 = self.r_res2(r2)
            r_pr[2] = r_pr[2] + res2
            out.insert(0, r_pr[2]); scales.insert(0, 2)
            i1 = self.r_iconv1(crop_cat(self.r_deconv1(r2), self._up(res2), r1))
            res1 = self.r_res1(i1)
            r_pr[1] = r_pr[1] + res1
            out.insert(0, r_pr[1]); scales.insert(0, 1)
            i0 = self.r_iconv0(crop_cat(self.r_deconv0(i1), self._up(res1), r0))
            r_pr[0] = r_pr[0] + self.r_res0(i0)
            out.insert(0, r_pr[0]); scales.insert(0, 0)
        if mode == "test":
            out[-1] = out[-1].clamp(self.delt, maxD)
        return scales, out
