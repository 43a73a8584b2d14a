"""DispNetC with the 1-D correlation layer on the sm_100a kernel — drop-in for the reference's
``dispnetcorr`` (models/dispnetcorr.py:13-134; BASELINE.json config 1).

Only the correlation (``Corr1d(kernel_size=1, stride=1, D=41)``, dispnetcorr.py:27,77) is on the hot
path; the 2-D encoder/decoder is a caller of it and stays stock PyTorch (cuDNN), exactly as in the
reference.  The module is generated from a layer table; parameter names are the reference's
(``conv1.0.weight``, ``deconv5.0.bias``, ``pr3.weight`` ...) so its ``state_dict`` loads unchanged, and
``forward(imL, imR, mode)`` returns ``(out_scale, out)`` with the 7-level pyramid ``[pr0 .. pr6]``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .corr1d import Corr1d

# name, Cin, Cout, kernel, stride          (dispnetcorr.py:25-37; every block is conv + bias + ReLU, no BatchNorm)
ENCODER = [("conv1", 3, 64, 7, 2), ("conv2", 64, 128, 5, 2), ("redir", 128, 64, 1, 1),
           ("conv3a", 64 + 41, 256, 5, 2), ("conv3b", 256, 256, 3, 1), ("conv4a", 256, 512, 3, 2), ("conv4b", 512, 512, 3, 1),
           ("conv5a", 512, 512, 3, 2), ("conv5b", 512, 512, 3, 1), ("conv6a", 512, 1024, 3, 2), ("conv6b", 1024, 1024, 3, 1)]
# level, deconv Cin->Cout, iconv Cin (= Cout + 1 + skip channels)   (dispnetcorr.py:40-62)
DECODER = [(5, 1024, 512, 1025), (4, 512, 256, 769), (3, 256, 128, 385), (2, 128, 64, 193), (1, 64, 32, 97)]


def _conv_relu(cin, cout, k, s):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, padding=(k - 1) // 2, bias=True), nn.ReLU(inplace=True))


def _deconv_relu(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, padding=1, output_padding=0, bias=True), nn.ReLU(inplace=True))


def crop_cat(*ts):
    """myCat2d (models/util_fun.py:7-19): concatenate along channels after cropping to the smallest H, W."""
    h = min(t.shape[2] for t in ts); w = min(t.shape[3] for t in ts)
    return torch.cat([t[:, :, :h, :w] for t in ts], dim=1)


class dispnetcorr(nn.Module):
    def __init__(self, maxdisparity=192, align_corners=True):
        super().__init__()
        self.name = "dispnetcorr"
        self.D = maxdisparity
        self.delt = 1e-6
        self.count_levels = 7
        self.align_corners = align_corners        # nn.Upsample(bilinear) of PyTorch <= 0.3
        for name, cin, cout, k, s in ENCODER:
            setattr(self, name, _conv_relu(cin, cout, k, s))
        self.corr = Corr1d(kernel_size=1, stride=1, D=41, simfun=None)
        self.pr6 = nn.Conv2d(1024, 1, 3, 1, 1)
        for lvl, cin, cout, icin in DECODER:
            setattr(self, "deconv%d" % lvl, _deconv_relu(cin, cout))
            setattr(self, "iconv%d" % lvl, _conv_relu(icin, cout, 3, 1))
            setattr(self, "pr%d" % lvl, nn.Conv2d(cout, 1, 3, 1, 1))
        for m in self.modules():                                   # net_init (util_conv.py:32-53)
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)))
        for lvl in range(1, 7):                                    # dispnetcorr.py:63-64
            getattr(self, "pr%d" % lvl).weight.data.mul_(0.1)

    def _up(self, x):
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=self.align_corners)

    def forward(self, imL, imR, mode="train"):
        assert imL.shape == imR.shape
        maxD = max(self.D, imL.shape[-1])
        c1L, c1R = self.conv1(imL), self.conv1(imR)
        c2L, c2R = self.conv2(c1L), self.conv2(c1R)
        x = torch.cat([self.corr(c2L, c2R), self.redir(c2L)], dim=1)          # the hot-path op (dispnetcorr.py:77)
        skips = {1: c1L, 2: c2L}
        for lvl in (3, 4, 5, 6):
            x = getattr(self, "conv%db" % lvl)(getattr(self, "conv%da" % lvl)(x))
            skips[lvl] = x
        out = [self.pr6(x)]
        for lvl, _, _, _ in DECODER:
            x = getattr(self, "iconv%d" % lvl)(crop_cat(getattr(self, "deconv%d" % lvl)(x), self._up(out[0]), skips[lvl]))
            out.insert(0, getattr(self, "pr%d" % lvl)(x))
        out.insert(0, self._up(out[0])[:, :, :imL.shape[-2], :imL.shape[-1]])
        if mode == "test":
            out[-1] = out[-1].clamp(self.delt, maxD)                          # sic: the coarsest level
        return list(range(7)), out
