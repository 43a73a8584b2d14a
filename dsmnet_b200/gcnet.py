"""GC-Net's 3-D path on the sm_100a kernels: concat volume (models/gcnet.py:131-135), the 19-layer
3-D encoder-decoder ``feature3d`` (:32-101, blocks from models/util_conv.py:150-179) and the
soft-argmin of MINUS the cost (:104-111).

``feature3d`` keeps the reference's constructor and parameter names (``l19.0.weight``, ``l19.0.bias``,
``l19.1.running_var``, ..., bare ``l37.weight``) so that the ``layer3d.*`` part of a reference
``state_dict`` loads unchanged.  Inference: eval-mode BatchNorm folded into the conv epilogue, everything
fused.  With gradients enabled (train mode) the same graph runs with autograd: convolutions forward /
backward on the sm_100a kernels, BatchNorm (batch statistics) + ReLU + skip adds (cropped where sizes are odd) on the
fused streaming kernels of ``train3d``.  The 2-D trunk ``feature2d`` (:14-29) is a caller of the path and stays
stock PyTorch in the reference; ``GCNetHotPath`` therefore starts from the two feature maps.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib
from .conv3d import FusedConv3d, conv_out_dims
from .cost_volume import concat_volume, concat_volume_padded
from .softargmin import softargmin
from .volume_layout import PaddedVolume
from .psmnet import PDL_VARIANT, CLS_SIDE_STREAM


def conv3d_bn(in_planes, out_planes, kernel_size=3, stride=1):
    """Parameter container with the reference's layout (util_conv.py:150-162); ReLU has no parameters."""
    return nn.Sequential(nn.Conv3d(in_planes, out_planes, kernel_size, stride, padding=(kernel_size - 1) // 2, bias=True),
                         nn.BatchNorm3d(out_planes), nn.ReLU(inplace=True))


def deconv3d_bn(in_planes, out_planes, kernel_size=3, stride=2, bn=True):
    """util_conv.py:164-179 (k3, s2, p1, output_padding 1).  The reference appends BatchNorm2d here (a bug that
    raises on 5-D input); BatchNorm3d has identical parameter names and shapes."""
    p = (kernel_size - 1) // 2
    op = stride - (kernel_size - 2 * p)
    conv = nn.ConvTranspose3d(in_planes, out_planes, kernel_size, stride, padding=p, output_padding=op, bias=True)
    if not bn:
        return conv
    return nn.Sequential(conv, nn.BatchNorm3d(out_planes), nn.ReLU(inplace=True))


class feature3d(nn.Module):
    def __init__(self, num_F=32):
        super().__init__()
        F_ = self.F = num_F
        self.l19 = conv3d_bn(F_ * 2, F_); self.l20 = conv3d_bn(F_, F_)
        self.l21 = conv3d_bn(F_ * 2, F_ * 2, stride=2); self.l22 = conv3d_bn(F_ * 2, F_ * 2); self.l23 = conv3d_bn(F_ * 2, F_ * 2)
        self.l24 = conv3d_bn(F_ * 2, F_ * 2, stride=2); self.l25 = conv3d_bn(F_ * 2, F_ * 2); self.l26 = conv3d_bn(F_ * 2, F_ * 2)
        self.l27 = conv3d_bn(F_ * 2, F_ * 2, stride=2); self.l28 = conv3d_bn(F_ * 2, F_ * 2); self.l29 = conv3d_bn(F_ * 2, F_ * 2)
        self.l30 = conv3d_bn(F_ * 2, F_ * 4, stride=2); self.l31 = conv3d_bn(F_ * 4, F_ * 4); self.l32 = conv3d_bn(F_ * 4, F_ * 4)
        self.l33 = deconv3d_bn(F_ * 4, F_ * 2); self.l34 = deconv3d_bn(F_ * 2, F_ * 2)
        self.l35 = deconv3d_bn(F_ * 2, F_ * 2); self.l36 = deconv3d_bn(F_ * 2, F_)
        self.l37 = deconv3d_bn(F_, 1, bn=False)
        self._plan = None
        self._side = {}
        self._plan_key = None
        self._ws: Dict[Tuple, dict] = {}

    # -- packed weights + folded BatchNorm, rebuilt when a parameter changes ------------------------
    def _wants_autograd(self, *inputs):
        if self.training:
            return True
        if torch.is_grad_enabled() and (any(t.requires_grad for t in inputs) or any(p.requires_grad for p in self.parameters())):
            if not getattr(self, "_warned_eval_grad", False):
                import warnings
                warnings.warn("dsmnet_b200: eval-mode forward with autograd enabled takes the differentiable (unfused) path; "
                              "wrap inference in torch.no_grad() for the fused kernels", stacklevel=3)
                self._warned_eval_grad = True
            return True
        return False

    def _get_plan(self, device):
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()) if t.dim() > 0)
        if self._plan is None or key != self._plan_key:
            plan = {}
            for name in ("l19", "l20", "l21", "l22", "l23", "l24", "l25", "l26", "l27", "l28", "l29", "l30", "l31", "l32",
                         "l33", "l34", "l35", "l36"):
                seq = getattr(self, name)
                tr = isinstance(seq[0], nn.ConvTranspose3d)
                plan[name] = FusedConv3d(seq[0].weight, seq[1], seq[0].bias, seq[0].stride[0], tr, 2, device, PDL_VARIANT)
            plan["l37"] = FusedConv3d(self.l37.weight, None, self.l37.bias, 2, True, 0, device, PDL_VARIANT)
            self._plan, self._plan_key = plan, key
        return self._plan

    def _workspace(self, B, D, H, W, device):
        """Activation buffers (zero rims written once).  Skip tensors fix the extent of the deconv outputs
        that are added to them (myAdd3d crop-to-min, util_fun.py:41-51)."""
        key = (B, D, H, W, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        dims = [(D, H, W)]
        for _ in range(4):
            dims.append(conv_out_dims(*dims[-1], 2, False))
        P = PaddedVolume.empty
        F_ = self.F
        ws = {"x19": P(B, F_, *dims[0], device), "x20": P(B, F_, *dims[0], device)}
        for lvl, (a, b, c) in ((1, ("x21", "x22", "x23")), (2, ("x24", "x25", "x26")), (3, ("x27", "x28", "x29"))):
            for n in (a, b, c):
                ws[n] = P(B, 2 * F_, *dims[lvl], device)
        for n in ("x30", "x31", "x32"):
            ws[n] = P(B, 4 * F_, *dims[4], device)
        # deconv outputs: min(2*in, skip) per axis
        def up(src, skip):
            return tuple(min(2 * s, k) for s, k in zip(src, skip))
        d33 = up(dims[4], dims[3]); d34 = up(d33, dims[2]); d35 = up(d34, dims[1]); d36 = up(d35, dims[0])
        ws["x33"] = P(B, 2 * F_, *d33, device); ws["x34"] = P(B, 2 * F_, *d34, device)
        ws["x35"] = P(B, 2 * F_, *d35, device); ws["x36"] = P(B, F_, *d36, device)
        ws["x37"] = torch.empty(B, 2 * d36[0], 2 * d36[1], 2 * d36[2], device=device, dtype=torch.float32)
        self._ws[key] = ws
        return ws

    def aggregate_train(self, vol: PaddedVolume) -> torch.Tensor:
        """The same graph with autograd (gcnet.py:65-101 under model.train()): convolutions fwd/bwd on the sm_100a
        kernels, BatchNorm with batch statistics + ReLU + (cropped) skip adds on the kernels of dsmnet_b200/train3d.py."""
        from . import train3d as T

        def g(name, x, residual=None):
            seq = getattr(self, name)
            return T.conv_bn_act(x, seq[0], seq[1], 2, residual)       # conv+bias -> BN -> ReLU, then the skip add

        x21 = g("l21", vol); x24 = g("l24", x21); x27 = g("l27", x24); x30 = g("l30", x27)
        x32 = g("l32", g("l31", x30))
        x29 = g("l29", g("l28", x27)); x33 = g("l33", x32, x29)
        x26 = g("l26", g("l25", x24)); x34 = g("l34", x33, x26)
        x23 = g("l23", g("l22", x21)); x35 = g("l35", x34, x23)
        x20 = g("l20", g("l19", vol)); x36 = g("l36", x35, x20)
        return T.conv_c1(x36, self.l37)

    def aggregate(self, vol: PaddedVolume) -> torch.Tensor:
        """x37 of gcnet.py:65-101 as fp32 [B, 2D, 2H, 2W]."""
        if self._wants_autograd(vol.data):
            return self.aggregate_train(vol)
        dev = vol.data.device
        p = self._get_plan(dev)
        ws = self._workspace(vol.B, vol.D, vol.H, vol.W, dev)

        def crop(skip: PaddedVolume, like: PaddedVolume) -> PaddedVolume:
            """the skip tensor cropped to the deconv output's extent (myAdd3d); a no-op for even sizes"""
            if (skip.D, skip.H, skip.W) == (like.D, like.H, like.W):
                return skip
            return PaddedVolume.from_ncdhw(skip.to_ncdhw()[:, :, :like.D, :like.H, :like.W])

        # the full-resolution skip branch l19 -> l20 (59 % of the flops) is needed only by l36: it runs on a second
        # stream beside the encoder-decoder chain, whose 1/8 .. 1/32 resolution layers cannot fill the machine
        side = None
        if CLS_SIDE_STREAM:
            main = torch.cuda.current_stream(dev)
            side = self._side.get(str(dev))
            if side is None:
                side = self._side[str(dev)] = torch.cuda.Stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                x20 = p["l20"](p["l19"](vol, ws["x19"]), ws["x20"])
        x21 = p["l21"](vol, ws["x21"]); x24 = p["l24"](x21, ws["x24"]); x27 = p["l27"](x24, ws["x27"]); x30 = p["l30"](x27, ws["x30"])
        x32 = p["l32"](p["l31"](x30, ws["x31"]), ws["x32"])
        x29 = p["l29"](p["l28"](x27, ws["x28"]), ws["x29"])
        x33 = p["l33"](x32, ws["x33"], residual=crop(x29, ws["x33"]))
        x26 = p["l26"](p["l25"](x24, ws["x25"]), ws["x26"])
        x34 = p["l34"](x33, ws["x34"], residual=crop(x26, ws["x34"]))
        x23 = p["l23"](p["l22"](x21, ws["x22"]), ws["x23"])
        x35 = p["l35"](x34, ws["x35"], residual=crop(x23, ws["x35"]))
        if side is None:
            x20 = p["l20"](p["l19"](vol, ws["x19"]), ws["x20"])
        else:
            main.wait_stream(side)
        x36 = p["l36"](x35, ws["x36"], residual=crop(x20, ws["x36"]))
        return p["l37"](x36, ws["x37"])

    def forward(self, x, mode="train"):
        """Reference signature (gcnet.py:65): x is the NCDHW fp32 cost volume; returns [B,1,2H,2W]."""
        x37 = self.aggregate(PaddedVolume.from_ncdhw(x))
        return softargmin(x37, -1.0).unsqueeze(1)


class GCNetHotPath(nn.Module):
    """(fL, fR) feature maps -> disparity [B,1,2h,2w]: the path of gcnet.forward after ``layer2d`` (gcnet.py:129-137)."""

    def __init__(self, maxdisparity=192):
        super().__init__()
        self.D = maxdisparity // 2            # gcnet.py:117 (Py2 integer division)
        self.layer3d = feature3d(32)

    def forward(self, fL, fR):
        if self.training or self.layer3d._wants_autograd(fL, fR):
            vol = concat_volume_padded(fL, fR, self.D, "gc")                 # differentiable, padded bf16 directly
        else:
            vol = concat_volume(fL, fR, self.D, "gc", padded_bf16=True)
        x37 = self.layer3d.aggregate(vol)
        return softargmin(x37, -1.0).unsqueeze(1)


# --------------------------------------------------------------------------------------------
# 2-D trunk and whole-model drop-in (gcnet.py:14-29, 113-137), reference parameter names.
# --------------------------------------------------------------------------------------------

class _BasicBlock2d(nn.Module):
    """util_conv.BasicBlock (models/util_conv.py:180-208): conv-bn-relu-conv-bn, + x, relu."""

    def __init__(self, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False); self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False); self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        out = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        return self.relu(out + x)


class feature2d(nn.Module):
    """gcnet.py:14-29.  Inference on CUDA: trunk2d.GCNetTrunkPlan (one launch per layer); otherwise stock PyTorch."""

    def __init__(self, num_F=32):
        super().__init__()
        self.F = num_F
        self.conv1 = nn.Sequential(nn.Conv2d(3, 32, 5, 2, 2, bias=True), nn.BatchNorm2d(32), nn.ReLU(inplace=True))
        self.block1 = nn.Sequential(*[_BasicBlock2d(32) for _ in range(8)])
        self.conv2 = nn.Conv2d(32, 32, 3, 1, 1)

    def forward(self, x):
        if x.is_cuda and not self.training and not torch.is_grad_enabled():
            from .trunk2d import GCNetTrunkPlan, cached_plan
            return cached_plan(self, GCNetTrunkPlan, x.device)(x)
        return self.conv2(self.block1(self.conv1(x)))


class gcnet(nn.Module):
    """Drop-in for the reference's gcnet (gcnet.py:113-137): forward(imL, imR, mode) -> ([0], [disparity [B,1,H,W]])."""

    def __init__(self, maxdisparity=192):
        super().__init__()
        self.name = "gcnet"
        self.D = maxdisparity // 2            # gcnet.py:117 (Py2 integer division)
        self.count_levels = 1
        self.layer2d = feature2d(32)
        self.layer3d = feature3d(32)

    def forward(self, imL, imR, mode="train"):
        assert imL.shape == imR.shape
        B = imL.size(0)
        if imL.is_cuda and not self.training and not torch.is_grad_enabled():
            f = self.layer2d(torch.cat((imL, imR), 0))
            fL, fR = f[:B], f[B:]
            vol = concat_volume(fL, fR, self.D, "gc", padded_bf16=True)
        else:
            fL, fR = self.layer2d(imL), self.layer2d(imR)
            vol = concat_volume_padded(fL, fR, self.D, "gc")
        x37 = self.layer3d.aggregate(vol)
        oL = softargmin(x37, -1.0).unsqueeze(1)[:, :, :imL.shape[-2], :imL.shape[-1]]
        return [0], [oL]
