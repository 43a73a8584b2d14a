"""PSMNet (stacked hourglass) with the cost-volume hot path on the sm_100a kernels.

Mirrors the reference's model surface (models/psmnet/stackhourglass.py, submodule.py): same
class names, constructor arguments and parameter / buffer names (``dres0.0.0.weight``,
``dres2.conv5.1.running_var`` ...) so a reference ``state_dict`` loads unchanged, and
``PSMNet(maxdisp)(left, right, mode)`` returns ``([0, 0, 0], [pred3, pred2, pred1])`` with
``pred*`` of shape (B, H, W) exactly like stackhourglass.py:168.

The 2-D feature extractor (a caller of the hot path) runs on the library's 2-D kernels for inference
(dsmnet_b200/trunk2d.py) and as stock PyTorch under autograd; the
path from the two feature maps on — concat volume (stackhourglass.py:124-133), dres0..classif3
(:135-149) and the three upsample+softmax+regression heads (:152-166) — runs as:
  concat_volume (padded NDHWC bf16; or, with DSM_FUSED_VOLUME=1, dres0.0 builds its volume tiles itself and the volume
  is never written)  ->  25 fused tcgen05 conv blocks  ->  3 fp32 Cout=1 convs  ->  3 fused upsample+soft-argmin kernels.
Inference: eval-mode BatchNorm folded into the conv epilogue, everything fused.  With gradients enabled
(train mode, or eval-mode with parameters/inputs that require grad) the same graph runs through
``aggregate_train``: convolutions forward/backward on the sm_100a kernels, BatchNorm (batch statistics or frozen) +
ReLU + skip adds on the fused streaming kernels of ``train3d`` (no stock-PyTorch ops on the volumes).
"""
from __future__ import annotations

import os

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .conv3d import FusedConv3d, conv_from_features, pack_feature_pair_nhwc
from .cost_volume import concat_volume, concat_volume_padded
from .softargmin import upsample_softargmin
from .volume_layout import PaddedVolume


def convbn_3d(in_planes, out_planes, kernel_size, stride, pad):
    """Parameter container with the reference's layout (submodule.py:16-19)."""
    return nn.Sequential(nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=stride, bias=False),
                         nn.BatchNorm3d(out_planes))


class hourglass(nn.Module):
    """Parameters as in stackhourglass.py:22-41; the forward is driven by PSMNetHotPath."""

    def __init__(self, inplanes):
        super().__init__()
        self.conv1 = nn.Sequential(convbn_3d(inplanes, inplanes * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = convbn_3d(inplanes * 2, inplanes * 2, 3, 1, 1)
        self.conv3 = nn.Sequential(convbn_3d(inplanes * 2, inplanes * 2, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(inplanes * 2, inplanes * 2, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = nn.Sequential(nn.ConvTranspose3d(inplanes * 2, inplanes * 2, kernel_size=3, padding=1,
                                                      output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm3d(inplanes * 2))
        self.conv6 = nn.Sequential(nn.ConvTranspose3d(inplanes * 2, inplanes, kernel_size=3, padding=1,
                                                      output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm3d(inplanes))


# dsm_conv3d_fwd_ex variant bit 7: programmatic dependent launch (DSM_NO_PDL=1 in the environment turns it off)
PDL_VARIANT = 0 if os.environ.get("DSM_NO_PDL") == "1" else 128
# classifier branches on a second stream next to the following hourglass (DSM_CLS_STREAM=0 turns it off)
CLS_SIDE_STREAM = os.environ.get("DSM_CLS_STREAM", "1") != "0"
# with the second stream: launch each head as soon as its cost exists (DSM_EARLY_HEADS=0: one stacked launch at the end)
EARLY_HEADS = os.environ.get("DSM_EARLY_HEADS", "1") != "0"
# dres0.0 reads the two feature maps and assembles its volume tiles itself (the 196 MB volume is never written);
# DSM_FUSED_VOLUME=0 materialises the volume instead (A/B: tools/ab_fused_volume.py)
FUSED_VOLUME = os.environ.get("DSM_FUSED_VOLUME", "1") != "0"


class _Plan:
    """Packed weights + folded BatchNorm of every 3-D layer, built once per parameter version."""

    def __init__(self, m: "PSMNetHotPath", device, variant=0):
        variant |= PDL_VARIANT        # consecutive layers: prologue of layer i+1 overlaps the tail of layer i
        def cb(seq, stride=1, transposed=False, relu=False):
            return FusedConv3d(seq[0].weight, seq[1], None, stride, transposed, relu, device, variant)

        self.dres0_0 = cb(m.dres0[0], relu=True)
        self.dres0_2 = cb(m.dres0[2], relu=True)
        self.dres1_0 = cb(m.dres1[0], relu=True)
        self.dres1_2 = cb(m.dres1[2])
        self.hg = []
        for h in (m.dres2, m.dres3, m.dres4):
            self.hg.append(dict(
                conv1=cb(h.conv1[0], stride=2, relu=True), conv2=cb(h.conv2, relu=True),
                conv3=cb(h.conv3[0], stride=2, relu=True), conv4=cb(h.conv4[0], relu=True),
                conv5=cb(h.conv5, transposed=True, relu=True), conv6=cb(h.conv6, transposed=True)))
        self.cls = []
        for c in (m.classif1, m.classif2, m.classif3):
            self.cls.append((cb(c[0], relu=True), FusedConv3d(c[2].weight, None, None, 1, False, False, device, variant)))


class PSMNetHotPath(nn.Module):
    """The north-star path as one module: (fL, fR) feature maps -> [pred3, pred2, pred1].

    Holds the 3-D parameters under the reference's names.  ``forward(fL, fR, out_hw)``."""

    def __init__(self, maxdisp=192, align_corners=True, variant=0):
        super().__init__()
        self.maxdisp = maxdisp
        self.align_corners = align_corners        # PyTorch<=0.3 F.upsample semantics
        self.variant = variant
        self._side = {}
        self.dres0 = nn.Sequential(convbn_3d(64, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                   convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                   convbn_3d(32, 32, 3, 1, 1))
        self.dres2 = hourglass(32)
        self.dres3 = hourglass(32)
        self.dres4 = hourglass(32)
        self.classif1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
        self.classif2 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
        self.classif3 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
        for m in self.modules():                  # init as stackhourglass.py:100-114
            if isinstance(m, nn.Conv3d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1); m.bias.data.zero_()
        self._plan: Optional[_Plan] = None
        self._plan_key = None
        self._ws: Dict[Tuple, dict] = {}

    # -- plan / workspace caches ------------------------------------------------------------
    _STACK = ("dres0", "dres1", "dres2", "dres3", "dres4", "classif1", "classif2", "classif3")

    def _stack_tensors(self):
        """Parameters and buffers of the 3-D stack only (a subclass may add a 2-D trunk with its own plan)."""
        return [t for n in self._STACK for m in (getattr(self, n),) for t in list(m.parameters()) + list(m.buffers())
                if t.dim() > 0]                                     # num_batches_tracked does not enter the plan

    def _wants_autograd(self, fL, fR):
        """The autograd graph is needed in train mode, or when gradients are on and something on the path asks for one.
        model.eval() without torch.no_grad() lands here too (parameters require grad by default): warn once, because that
        path computes batch-norm-frozen training kernels and saves activations — inference should run under no_grad."""
        if self.training:
            return True
        if torch.is_grad_enabled() and (fL.requires_grad or fR.requires_grad or any(t.requires_grad for t in self._stack_tensors())):
            if not getattr(self, "_warned_eval_grad", False):
                import warnings
                warnings.warn("dsmnet_b200: eval-mode forward with autograd enabled takes the differentiable (unfused) path; "
                              "wrap inference in torch.no_grad() for the fused kernels", stacklevel=3)
                self._warned_eval_grad = True
            return True
        return False

    def _get_plan(self, device):
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in self._stack_tensors())
        if self._plan is None or key != self._plan_key:
            self._plan = _Plan(self, device, self.variant)
            self._plan_key = key
        return self._plan

    def _workspace(self, B, D, H, W, device):
        """Activation buffers for one problem size (zero rims written once, then reused)."""
        key = (B, D, H, W, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        def half(n): return (n - 1) // 2 + 1
        D2, H2, W2 = half(D), half(H), half(W)
        D4, H4, W4 = half(D2), half(H2), half(W2)
        P = PaddedVolume.empty
        ws = dict(
            vol=None,                 # only the unfused route (DSM_FUSED_VOLUME=0) materialises the volume
            a=P(B, 32, D, H, W, device), c0=P(B, 32, D, H, W, device), t=P(B, 32, D, H, W, device),
            cost0=P(B, 32, D, H, W, device),
            out=[P(B, 32, D, H, W, device) for _ in range(3)],
            h1=P(B, 64, D2, H2, W2, device),
            pre=[P(B, 64, D2, H2, W2, device) for _ in range(3)],
            post=[P(B, 64, D2, H2, W2, device) for _ in range(3)],
            h3=P(B, 64, D4, H4, W4, device), h4=P(B, 64, D4, H4, W4, device),
            # the three classifier outputs in head order (cost3, cost2, cost1): one head launch reads all of them
            cost_all=torch.empty(3, B, D, H, W, device=device, dtype=torch.float32),
        )
        ws["cost"] = [ws["cost_all"][2 - i] for i in range(3)]      # cost[i] = cost_{i+1}
        self._ws[key] = ws
        return ws

    # -- the path -----------------------------------------------------------------------------
    def aggregate_train(self, fL, fR):
        """The same graph with autograd (stackhourglass.py:123-149): convolutions fwd/bwd on the sm_100a kernels,
        BatchNorm (batch statistics in train mode) + ReLU + skip adds on the fused kernels of dsmnet_b200/train3d.py."""
        from . import train3d as T
        D = self.maxdisp // 4
        vol = concat_volume_padded(fL, fR, D, "psm")      # padded bf16 directly; backward sums over d into NCHW fp32

        def cb(seq, x, relu=0, residual=None):
            return T.conv_bn_act(x, seq[0], seq[1], relu, residual)

        c0 = cb(self.dres0[0], vol, 1)
        c0 = cb(self.dres0[2], c0, 1)
        t = cb(self.dres1[0], c0, 1)
        cost0 = cb(self.dres1[2], t, 0, c0)

        def hg(h, x, presqu, postsqu):
            out = cb(h.conv1[0], x, 1)
            pre = cb(h.conv2, out, 1, postsqu)
            out = cb(h.conv3[0], pre, 1)
            out = cb(h.conv4[0], out, 1)
            post = cb(h.conv5, out, 1, presqu if presqu is not None else pre)
            return cb(h.conv6, post, 0, cost0), pre, post            # "+ cost0" of :139,142,145

        out1, pre1, post1 = hg(self.dres2, cost0, None, None)
        out2, pre2, post2 = hg(self.dres3, out1, pre1, post1)
        out3, pre3, post3 = hg(self.dres4, out2, pre1, post2)        # pre1, as in :144
        costs, prev = [], None
        for c, x in ((self.classif1, out1), (self.classif2, out2), (self.classif3, out3)):
            cur = T.conv_c1(cb(c[0], x, 1), c[2])
            prev = cur if prev is None else cur + prev               # :148-149
            costs.append(prev)
        return costs

    def aggregate(self, fL, fR, head=None):
        """concat volume + dres0..classif3 -> (cost1, cost2, cost3), fp32 [B, D/4, H/4, W/4] each.
        `head(i, cost)` (inference with the second stream only): called on the classifier stream right after cost{i+1}
        exists, so that the heads of cost1 / cost2 run beside the later hourglasses instead of after them; returns
        True if it was used (the caller then skips its own head launch)."""
        if self._wants_autograd(fL, fR):
            return self.aggregate_train(fL, fR)
        nhwc = fL.dtype == torch.bfloat16                 # zero-rimmed bf16 NHWC [B, H+2, W+2, 32] from the 2-D trunk's last layer
        if nhwc:
            B, H, W, C = fL.shape[0], fL.shape[1] - 2, fL.shape[2] - 2, fL.shape[3]
        else:
            B, C, H, W = fL.shape
        D = self.maxdisp // 4
        plan = self._get_plan(fL.device)
        ws = self._workspace(B, D, H, W, fL.device)
        if nhwc and not (FUSED_VOLUME and C == 32):
            fL = fL[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().contiguous(); fR = fR[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().contiguous()
            nhwc = False
        if FUSED_VOLUME and C == 32:
            # the 196 MB volume is never written: dres0.0 builds its operand tiles from the (L2-resident) feature maps
            hL, hR = (fL, fR) if nhwc else pack_feature_pair_nhwc(fL, fR)
            conv_from_features(plan.dres0_0, hL, hR, D, "psm", ws["a"])
        else:
            if ws["vol"] is None:
                ws["vol"] = PaddedVolume.empty(B, 64, D, H, W, fL.device, zero_rim=False)
            vol = concat_volume(fL, fR, D, "psm", padded_bf16=True, out=ws["vol"])
            plan.dres0_0(vol, ws["a"])
        plan.dres0_2(ws["a"], ws["c0"])
        plan.dres1_0(ws["c0"], ws["t"])
        cost0 = plan.dres1_2(ws["t"], ws["cost0"], residual=ws["c0"])            # :136
        x = cost0
        pre_prev = post_prev = None
        pre1 = None
        costs, prev = [], None
        side = main = None
        if CLS_SIDE_STREAM:
            main = torch.cuda.current_stream(fL.device)
            side = self._side.get(str(fL.device))
            if side is None:
                side = self._side[str(fL.device)] = torch.cuda.Stream(fL.device)
            if "tc" not in ws:
                ws["tc"] = [PaddedVolume.empty(B, 32, D, H, W, fL.device) for _ in range(3)]
        for i, hg in enumerate(plan.hg):
            presqu = None if i == 0 else pre1                                   # :141,144 (pre1 both times)
            postsqu = post_prev
            hg["conv1"](x, ws["h1"])
            pre = hg["conv2"](ws["h1"], ws["pre"][i], residual=postsqu)         # :46-50
            hg["conv3"](pre, ws["h3"])
            hg["conv4"](ws["h3"], ws["h4"])
            post = hg["conv5"](ws["h4"], ws["post"][i], residual=presqu if presqu is not None else pre)   # :55-58
            x = hg["conv6"](post, ws["out"][i], residual=cost0)                 # :60 then "+ cost0" (:139,142,145)
            if i == 0:
                pre1 = pre
            post_prev = post
            if side is not None:
                # classif{i+1} needs only out{i+1}: it runs on a second stream next to the following hourglass, whose
                # half / quarter resolution layers leave SMs idle (partial waves, pipeline fill and drain)
                ev = torch.cuda.Event(); ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    c0, c2 = plan.cls[i]
                    c0(ws["out"][i], ws["tc"][i])
                    prev = c2(ws["tc"][i], ws["cost"][i], residual=prev)       # :147-149 cumulative adds
                    costs.append(prev)
                    if head is not None:
                        head(i, prev)
        if side is not None:
            main.wait_stream(side)
            return costs
        for i, (c0, c2) in enumerate(plan.cls):
            c0(ws["out"][i], ws["t"])
            prev = c2(ws["t"], ws["cost"][i], residual=prev)                    # :147-149 cumulative adds
            costs.append(prev)
        return costs

    def host_pipeline(self, B, h, w, out_hw, device=None, use_graph=True):
        """Streaming host-buffer entry point: a `pipeline.HostPipeline` whose step takes two pinned host feature maps
        [B, 32, h, w] fp32 and yields the three disparity maps [B, H, W] in pinned host memory (H2D, CUDA-graph replay of
        the path, D2H; double-buffered).  The module must be in eval mode on a CUDA device."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        key = ("hot", B, h, w, tuple(out_hw), str(device), use_graph)
        ex = [torch.zeros(B, 32, h, w, device=device) for _ in range(2)]
        return _host_pipeline(self, key, lambda a, b: PSMNetHotPath.forward(self, a, b, tuple(out_hw)), ex, use_graph)

    def infer_host(self, fL_host, fR_host, out_hw):
        """One synchronous call with HOST feature maps -> [pred3, pred2, pred1] in pinned host memory."""
        B, _, h, w = fL_host.shape
        return self.host_pipeline(B, h, w, out_hw).run(fL_host, fR_host)

    def forward(self, fL, fR, out_hw):
        train = self._wants_autograd(fL, fR)
        if fL.dtype == torch.bfloat16 and train:
            raise _lib.DsmError("bf16 NHWC feature maps are an inference-only input of the hot path")
        if CLS_SIDE_STREAM and EARLY_HEADS and not train:
            # each head is launched on the classifier stream as soon as its cost exists: only the head of cost3 is left
            # on the critical path (one stacked launch of all three after the last classifier cost ~95 us more)
            B = fL.shape[0]
            preds = torch.empty(3, B, out_hw[0], out_hw[1], device=fL.device, dtype=torch.float32)
            size = (self.maxdisp, out_hw[0], out_hw[1])

            def head(i, cost):
                upsample_softargmin(cost, size, self.align_corners, out=preds[2 - i])      # preds = (pred3, pred2, pred1)
            self.aggregate(fL, fR, head)
            return [preds[0], preds[1], preds[2]]
        c1, c2, c3 = self.aggregate(fL, fR)
        if c1.requires_grad:
            # differentiable heads (stackhourglass.py:152-166): the fused upsample + soft-argmin kernels, forward and
            # backward, one launch each for the three stacked costs
            B = c1.shape[0]
            preds = upsample_softargmin(torch.cat((c3, c2, c1), 0), (self.maxdisp, out_hw[0], out_hw[1]), self.align_corners)
            return [preds[:B], preds[B:2 * B], preds[2 * B:]]
        B, H, W = (fL.shape[0], fL.shape[1] - 2, fL.shape[2] - 2) if fL.dtype == torch.bfloat16 else (fL.shape[0], fL.shape[2], fL.shape[3])
        ws = self._workspace(B, self.maxdisp // 4, H, W, fL.device)
        size = (self.maxdisp, out_hw[0], out_hw[1])
        # heads of stackhourglass.py:152-166 for (cost3, cost2, cost1) in ONE launch over the stacked costs
        stacked = ws["cost_all"].view(3 * B, *ws["cost_all"].shape[2:])
        preds = upsample_softargmin(stacked, size, self.align_corners).view(3, B, out_hw[0], out_hw[1])
        return [preds[0], preds[1], preds[2]]


def _host_pipeline(owner, key, step_fn, examples, use_graph=True):
    pipes = owner.__dict__.setdefault("_host_pipes", {})
    p = pipes.get(key)
    if p is None:
        from .pipeline import HostPipeline
        p = pipes[key] = HostPipeline(step_fn, examples, use_graph)
    return p


# --------------------------------------------------------------------------------------------
# 2-D feature extractor (caller of the hot path).  Same parameter names as the reference's
# feature_extraction (submodule.py:65-140) so that checkpoints load.  Inference on CUDA runs on the
# library's kernels (trunk2d.PSMNetTrunkPlan); with autograd or on CPU the stock-PyTorch graph below runs.
# --------------------------------------------------------------------------------------------

def convbn(in_planes, out_planes, kernel_size, stride, pad, dilation):
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                                   padding=dilation, dilation=dilation, bias=False),   # sic: `pad` unused (submodule.py:12)
                         nn.BatchNorm2d(out_planes))


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
        self.downsample = downsample

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        return out + (x if self.downsample is None else self.downsample(x))


class feature_extraction(nn.Module):
    def __init__(self, align_corners=True):
        super().__init__()
        self.align_corners = align_corners
        self.inplanes = 32
        self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), nn.ReLU(inplace=True),
                                       convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                       convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True))
        self.layer1 = self._make_layer(32, 3, 1, 1, 1)
        self.layer2 = self._make_layer(64, 16, 2, 1, 1)
        self.layer3 = self._make_layer(128, 3, 1, 1, 1)
        self.layer4 = self._make_layer(128, 3, 1, 1, 2)
        for i, k in enumerate((64, 32, 16, 8), start=1):
            setattr(self, "branch%d" % i, nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)),
                                                        convbn(128, 32, 1, 1, 0, 1), nn.ReLU(inplace=True)))
        self.lastconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, 32, kernel_size=1, padding=0, stride=1, bias=False))

    def _make_layer(self, planes, blocks, stride, pad, dilation):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample, pad, dilation)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x, nhwc_bf16=False):
        """`nhwc_bf16` (CUDA inference only): return the feature map as zero-rimmed bf16 [B, H/4+2, W/4+2, 32], the operand
        layout of the fused volume convolution, instead of the reference's fp32 NCHW (the tensor is the plan's own
        workspace: consume it before the next call)."""
        if x.is_cuda and not self.training and not torch.is_grad_enabled():
            # inference: every layer on the library's kernels (dsmnet_b200/trunk2d.py)
            from .trunk2d import PSMNetTrunkPlan, cached_plan
            return cached_plan(self, PSMNetTrunkPlan, x.device)(x, nhwc_bf16)
        if nhwc_bf16:
            raise _lib.DsmError("feature_extraction: the bf16 NHWC output exists on the CUDA inference path only")
        out = self.layer1(self.firstconv(x))
        raw = self.layer2(out)
        skip = self.layer4(self.layer3(raw))
        hw = (skip.size(2), skip.size(3))
        br = [F.interpolate(getattr(self, "branch%d" % i)(skip), hw, mode="bilinear", align_corners=self.align_corners)
              for i in (1, 2, 3, 4)]
        feat = torch.cat((raw, skip, br[3], br[2], br[1], br[0]), 1)
        return self.lastconv(feat)


class PSMNet(PSMNetHotPath):
    """Drop-in for the reference's PSMNet (stackhourglass.py:64-168)."""

    def __init__(self, maxdisp=192, align_corners=True):
        super().__init__(maxdisp, align_corners)
        self.name = "psmnet"
        self.count_levels = 1
        self.feature_extraction = feature_extraction(align_corners)
        for m in self.feature_extraction.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1); m.bias.data.zero_()

    def image_pipeline(self, B, H, W, device=None, use_graph=True):
        """Whole-model host-buffer entry point: step(left, right) with pinned host images [B, 3, H, W] fp32 -> the three
        disparity maps [B, H, W] in pinned host memory (trunk + hot path as one CUDA graph, copies double-buffered)."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        key = ("img", B, H, W, str(device), use_graph)
        ex = [torch.zeros(B, 3, H, W, device=device) for _ in range(2)]
        return _host_pipeline(self, key, lambda a, b: self.forward(a, b, "test")[1], ex, use_graph)

    def infer_host_images(self, left_host, right_host):
        B, _, H, W = left_host.shape
        return self.image_pipeline(B, H, W).run(left_host, right_host)

    def forward(self, left, right, mode="train"):
        B = left.size(0)
        if left.is_cuda and not self.training and not torch.is_grad_enabled():
            # both images as one batch through the trunk; its last layer writes the bf16 NHWC maps dres0.0 reads
            fea = self.feature_extraction(torch.cat((left, right), 0), nhwc_bf16=FUSED_VOLUME)
            refimg_fea, targetimg_fea = fea[:B], fea[B:]
        else:
            refimg_fea = self.feature_extraction(left).float()
            targetimg_fea = self.feature_extraction(right).float()
        preds = PSMNetHotPath.forward(self, refimg_fea, targetimg_fea, (left.size(2), left.size(3)))
        return [0, 0, 0], preds
