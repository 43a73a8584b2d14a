"""CPU oracle for the stereo cost-volume hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in closed form on CPU (torch CPU fp32/fp64 tensors + numpy), of what the reference
sunshinnnn/DSMnet computes on the path this repo accelerates.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package; nothing under ``dsmnet_b200/`` does.

Parity status: the reference ships NO golden vectors or tests for this path, so
the oracle is pinned against outputs of the reference's own code, executed in the build
container through ``oracle/refshim.py`` and frozen as fixtures by ``tests/golden/make_golden.py``
(``tests/test_oracle_golden.py`` re-checks every fixture on CPU; when ``/root/reference`` is
present ``tests/test_oracle_vs_reference.py`` re-runs the reference live).

The arithmetic of conv3d / conv_transpose3d / batch-norm / grid_sample / trilinear upsample /
softmax lives in a third-party dependency that is not vendored in the reference: PyTorch
(pinned by the reference only as "PyTorch 0.3.0+", README.md:25-26; this container has
torch 2.11.0+cu128).  For those the oracle calls the same torch CPU functional with the
PyTorch<=0.3 semantics pinned explicitly (align_corners=True), which is the reference's own call.

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# op 1 — Corr1d   (models/util_conv.py:56-86)
# ------------------------------------------------------------------------------------------

def corr1d(fL: torch.Tensor, fR: torch.Tensor, D: int, stride: int = 1, kernel_size: int = 1) -> torch.Tensor:
    """out[b,d,y,x] = sum_c fL[b,c,y,x]*fR[b,c,y,x-d*stride] for x>=d*stride, d<W; else 0.

    Restates Corr1d.forward (util_conv.py:71-86) with simfun_default (:68-69): no 1/C
    normalisation; the loop stops at d >= W (:79); kernel_size>1 appends a k x k stride-1
    zero-padded average pool with count_include_pad (:82-85)."""
    B, C, H, W = fL.shape
    out = fL.new_zeros(B, D, H, W)
    for d in range(min(D, W)):
        s = d * stride
        if s >= W:
            continue
        out[:, d, :, s:] = (fL[:, :, :, s:] * fR[:, :, :, : W - s]).sum(dim=1)
    if kernel_size > 1:
        assert kernel_size % 2 == 1
        out = F.avg_pool2d(out, kernel_size, stride=1, padding=kernel_size // 2)
    return out


def corr1d_grads(g: torch.Tensor, fL: torch.Tensor, fR: torch.Tensor, stride: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed-form backward of corr1d without pooling:
    gL[b,c,y,x] = sum_d g[b,d,y,x]*fR[b,c,y,x-d*s];  gR[b,c,y,x'] = sum_d g[b,d,y,x'+d*s]*fL[b,c,y,x'+d*s]."""
    B, C, H, W = fL.shape
    D = g.shape[1]
    gL = torch.zeros_like(fL)
    gR = torch.zeros_like(fR)
    for d in range(min(D, W)):
        s = d * stride
        if s >= W:
            continue
        gd = g[:, d : d + 1, :, s:]
        gL[:, :, :, s:] += gd * fR[:, :, :, : W - s]
        gR[:, :, :, : W - s] += gd * fL[:, :, :, s:]
    return gL, gR


# ------------------------------------------------------------------------------------------
# op 2 — concatenation cost volume
# ------------------------------------------------------------------------------------------

def concat_volume(fL: torch.Tensor, fR: torch.Tensor, D: int, mode: str = "psm") -> torch.Tensor:
    """NCDHW volume [B,2C,D,H,W].

    mode "psm"      : models/psmnet/stackhourglass.py:124-133 — both halves zero for x<d.
    mode "gc"       : models/gcnet.py:131-135 — left half copied for every x, right half shifted.
    mode "gc_right" : models/gcnet.py:157,159,163-164 (xR) — fR copied, fL shifted the other way."""
    B, C, H, W = fL.shape
    vol = fL.new_zeros(B, 2 * C, D, H, W)
    for d in range(D):
        if mode == "psm":
            if d < W:
                vol[:, :C, d, :, d:] = fL[:, :, :, d:]
                vol[:, C:, d, :, d:] = fR[:, :, :, : W - d]
        elif mode == "gc":
            vol[:, :C, d] = fL
            if d < W:
                vol[:, C:, d, :, d:] = fR[:, :, :, : W - d]
        elif mode == "gc_right":
            vol[:, :C, d] = fR
            if d < W:
                vol[:, C:, d, :, : W - d] = fL[:, :, :, d:]
        else:
            raise ValueError(mode)
    return vol


def concat_volume_grads(g: torch.Tensor, D: int, mode: str = "psm") -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed-form backward: returns (gL, gR), each [B,C,H,W]."""
    B, C2, D_, H, W = g.shape
    C = C2 // 2
    gA = g.new_zeros(B, C, H, W)   # gradient of the tensor copied into the first half
    gB = g.new_zeros(B, C, H, W)   # gradient of the shifted tensor
    for d in range(D):
        if mode == "psm":
            if d < W:
                gA[:, :, :, d:] += g[:, :C, d, :, d:]
                gB[:, :, :, : W - d] += g[:, C:, d, :, d:]
        elif mode == "gc":
            gA += g[:, :C, d]
            if d < W:
                gB[:, :, :, : W - d] += g[:, C:, d, :, d:]
        elif mode == "gc_right":
            gA += g[:, :C, d]
            if d < W:
                gB[:, :, :, d:] += g[:, C:, d, :, : W - d]
        else:
            raise ValueError(mode)
    if mode == "gc_right":
        return gB, gA          # first half was fR, shifted one was fL
    return gA, gB


# ------------------------------------------------------------------------------------------
# op 3 — 3-D convolution blocks
# ------------------------------------------------------------------------------------------

def fold_bn(weight_shape_cout: int, bn: Optional[Dict[str, torch.Tensor]], bias: Optional[torch.Tensor],
            eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm3d (+conv bias) as a per-channel affine:
    scale = gamma/sqrt(var+eps), shift = beta - mean*scale (+ bias*scale)."""
    if bn is None:
        scale = torch.ones(weight_shape_cout)
        shift = torch.zeros(weight_shape_cout)
    else:
        scale = bn["weight"] / torch.sqrt(bn["running_var"] + eps)
        shift = bn["bias"] - bn["running_mean"] * scale
    if bias is not None:
        shift = shift + bias * scale
    return scale.float(), shift.float()


def crop_add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """myadd_3d / myAdd3d crop-to-min add (stackhourglass.py:10-20, util_fun.py:41-51)."""
    d = min(a.shape[2], b.shape[2]); h = min(a.shape[3], b.shape[3]); w = min(a.shape[4], b.shape[4])
    return a[:, :, :d, :h, :w] + b[:, :, :d, :h, :w]


def conv3d_block(x: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
                 shift: Optional[torch.Tensor] = None, stride: int = 1, transposed: bool = False,
                 residual: Optional[torch.Tensor] = None, relu=False,
                 operand_dtype: Optional[torch.dtype] = None,
                 storage_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Conv3d(k3,p1,stride) or ConvTranspose3d(k3,s2,p1,output_padding=1), then the folded
    BatchNorm affine, then optional crop-add of `residual`, then optional ReLU (`relu=2`: ReLU first, then the add).

    Restates convbn_3d (submodule.py:16-19), the hourglass wiring (stackhourglass.py:26-41,45-60)
    and conv3d_bn/deconv3d_bn (util_conv.py:150-179).  `operand_dtype=torch.bfloat16` rounds the
    conv operands (not the accumulation) the way the tensor-core kernel does; `storage_dtype`
    additionally rounds the block's result the way the kernel stores activations (so a following
    residual add reads the rounded value).  Both None = the reference's fp32 arithmetic."""
    if operand_dtype is not None:
        x = x.to(operand_dtype).float()
        weight = weight.to(operand_dtype).float()
    if transposed:
        y = F.conv_transpose3d(x, weight, None, stride=2, padding=1, output_padding=1)
    else:
        y = F.conv3d(x, weight, None, stride=stride, padding=1)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1, 1)
    if relu == 2:          # GC-Net: the block is activated first, the skip is added afterwards (gcnet.py:78-96)
        y = F.relu(y)
    if residual is not None:
        y = crop_add(y, residual)
    if relu == 1 or relu is True:
        y = F.relu(y)
    if storage_dtype is not None:
        y = y.to(storage_dtype).float()
    return y


# ------------------------------------------------------------------------------------------
# op 4 — soft-argmin
# ------------------------------------------------------------------------------------------

def disparity_regression(prob: torch.Tensor) -> torch.Tensor:
    """disparityregression.forward (submodule.py:60-63): sum_d d*prob[b,d,y,x] -> [B,H,W]."""
    disp = torch.arange(0, prob.shape[1], dtype=prob.dtype)
    return prob.permute(0, 2, 3, 1).matmul(disp)


def softargmin(cost: torch.Tensor, sign: float = 1.0) -> torch.Tensor:
    """softmax over dim 1 of sign*cost, then regression.  PSMNet: stackhourglass.py:155-157
    (F.softmax without dim on a 4-D tensor = dim 1), sign=+1; GC-Net: gcnet.py:104-109
    (Softmax2d of MINUS the cost), sign=-1.  Returns [B,H,W]."""
    return disparity_regression(F.softmax(sign * cost, dim=1))


def upsample_softargmin(cost_lr: torch.Tensor, size: Sequence[int], align_corners: bool = True) -> torch.Tensor:
    """PSMNet head (stackhourglass.py:152-166): F.upsample(cost[B,1,Dl,Hl,Wl], [D,H,W], 'trilinear')
    -> squeeze(1) -> softmax(dim 1) -> regression.  `align_corners=True` is what F.upsample did in
    the PyTorch the reference targets (<= 0.3.1)."""
    up = F.interpolate(cost_lr.unsqueeze(1), size=list(size), mode="trilinear", align_corners=align_corners)
    return softargmin(up.squeeze(1), 1.0)


# ------------------------------------------------------------------------------------------
# op 5 — imwrap
# ------------------------------------------------------------------------------------------

def imwrap_rowcol(h0: int, w0: int, h: int, w: int, LeftTop=(0, 0), scale_factor=1) -> Tuple[torch.Tensor, torch.Tensor]:
    """The two torch.linspace vectors of imwrap.py:50-58, computed exactly as there (python
    floats -> torch.linspace fp32)."""
    x, y = LeftTop
    x = x * 2.0 / (w0 - 1) - 1
    y = y * 2.0 / (h0 - 1) - 1
    x1 = x + (w - 1) * scale_factor * 2.0 / (w0 - 1)
    y1 = y + (h - 1) * scale_factor * 2.0 / (h0 - 1)
    return torch.linspace(x, x1, w), torch.linspace(y, y1, h)


def imwrap(im_src: torch.Tensor, disp: torch.Tensor, fliplr: bool = False, LeftTop=(0, 0), scale_factor=1,
           delt: float = 0.0) -> torch.Tensor:
    """imwrap_BCHW (utils/imwrap.py:37-72) with `delt` explicit instead of drawn at :70.
    grid_sample is pinned to bilinear / zeros / align_corners=True (the behaviour the grid of
    :51-54 is built for)."""
    bn, _, h0, w0 = im_src.shape
    bn, c, h, w = disp.shape
    assert c == 1 and min(h, w, h0, w0) > 1
    row, col = imwrap_rowcol(h0, w0, h, w, LeftTop, scale_factor)
    grid = torch.zeros(bn, h, w, 2)
    grid[..., 0] = row.view(1, 1, w)
    grid[..., 1] = col.view(1, h, 1)
    grid = grid.type_as(im_src)
    k = -1.0 if fliplr else 1
    gx = k * (grid[:, :, :, 0] - disp.squeeze(1) * 2.0 / (w0 - 1))
    grid = torch.stack([gx, grid[:, :, :, 1]], dim=3)
    return F.grid_sample(im_src + delt, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def imwrap_closed_form(im_src: np.ndarray, disp: np.ndarray, row: np.ndarray, col: np.ndarray, fliplr: bool,
                       delt: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Independent numpy restatement of the same op in fp32; returns
    (out, x0, y0) where x0/y0 are the integer north-west source indices (the 'warp indexing'
    that must match bit-exactly)."""
    f = np.float32
    B, C, H0, W0 = im_src.shape
    _, _, H, W = disp.shape
    k = f(-1.0) if fliplr else f(1.0)
    q = (disp[:, 0].astype(f) * f(2.0)) / f(W0 - 1)
    gx = k * (row.astype(f)[None, None, :] - q)
    gy = np.broadcast_to(col.astype(f)[None, :, None], gx.shape)
    ix = ((gx + f(1)) / f(2)) * f(W0 - 1)
    iy = ((gy + f(1)) / f(2)) * f(H0 - 1)
    x0 = np.floor(ix); y0 = np.floor(iy)
    tx = (ix - x0).astype(f); ty = (iy - y0).astype(f)
    ex = f(1) - tx; ey = f(1) - ty
    x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)
    out = np.zeros((B, C, H, W), dtype=f)
    src = im_src.astype(f) + f(delt)
    bidx = np.arange(B)[:, None, None]
    for (dx, dy, wgt) in ((0, 0, ex * ey), (1, 0, tx * ey), (0, 1, ex * ty), (1, 1, tx * ty)):
        xx = x0 + dx; yy = y0 + dy
        ok = (xx >= 0) & (xx < W0) & (yy >= 0) & (yy < H0)
        xc = np.clip(xx, 0, W0 - 1); yc = np.clip(yy, 0, H0 - 1)
        v = src[bidx, :, yc, xc]                      # [B,H,W,C]
        out += np.moveaxis(v * (wgt * ok)[..., None], 3, 1).astype(f)
    return out, x0, y0


# ------------------------------------------------------------------------------------------
# the north-star path: PSMNet cost volume -> stacked hourglass -> three soft-argmin heads
# ------------------------------------------------------------------------------------------

def _bn(params: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k: params[prefix + "." + k] for k in ("weight", "bias", "running_mean", "running_var")}


def _convbn(params, prefix, x, stride=1, transposed=False, residual=None, relu=False, operand_dtype=None):
    """`prefix` names an nn.Sequential(conv, bn) as in convbn_3d (submodule.py:16-19).
    `operand_dtype` may be a (operand, storage) pair."""
    w = params[prefix + ".0.weight"]
    cout = w.shape[1] if transposed else w.shape[0]
    scale, shift = fold_bn(cout, _bn(params, prefix + ".1"), None)
    od, sd = operand_dtype if isinstance(operand_dtype, tuple) else (operand_dtype, None)
    return conv3d_block(x, w, scale, shift, stride, transposed, residual, relu, od, sd)


def _hourglass(params, p, x, presqu, postsqu, operand_dtype=None, add_after=None):
    """hourglass.forward (stackhourglass.py:43-62); `add_after` is the caller's myadd_3d(out, cost0)."""
    out = _convbn(params, p + ".conv1.0", x, stride=2, relu=True, operand_dtype=operand_dtype)
    pre = _convbn(params, p + ".conv2", out, residual=postsqu, relu=True, operand_dtype=operand_dtype)
    out = _convbn(params, p + ".conv3.0", pre, stride=2, relu=True, operand_dtype=operand_dtype)
    out = _convbn(params, p + ".conv4.0", out, relu=True, operand_dtype=operand_dtype)
    post = _convbn(params, p + ".conv5", out, transposed=True, residual=presqu if presqu is not None else pre,
                   relu=True, operand_dtype=operand_dtype)
    out = _convbn(params, p + ".conv6", post, transposed=True, residual=add_after, operand_dtype=operand_dtype)
    return out, pre, post


def psmnet_aggregate(params: Dict[str, torch.Tensor], cost: torch.Tensor, operand_dtype=None):
    """dres0 .. classif3 of PSMNet.forward (stackhourglass.py:135-149) with eval-mode BatchNorm.
    Returns the three low-resolution costs (cost1, cost2, cost3), each [B,1,D/4,H/4,W/4]."""
    od = operand_dtype
    c0 = _convbn(params, "dres0.0", cost, relu=True, operand_dtype=od)
    c0 = _convbn(params, "dres0.2", c0, relu=True, operand_dtype=od)
    t = _convbn(params, "dres1.0", c0, relu=True, operand_dtype=od)
    cost0 = _convbn(params, "dres1.2", t, residual=c0, operand_dtype=od)

    # the "+ cost0" of :139,142,145 is the residual of conv6 (so that a storage dtype rounds once)
    out1, pre1, post1 = _hourglass(params, "dres2", cost0, None, None, od, cost0)
    out2, pre2, post2 = _hourglass(params, "dres3", out1, pre1, post1, od, cost0)
    out3, pre3, post3 = _hourglass(params, "dres4", out2, pre1, post2, od, cost0)   # NB: pre1, as in :144

    def classif(p, x):
        t = _convbn(params, p + ".0", x, relu=True, operand_dtype=od)
        return conv3d_block(t, params[p + ".2.weight"], operand_dtype=od[0] if isinstance(od, tuple) else od)

    cost1 = classif("classif1", out1)
    cost2 = classif("classif2", out2) + cost1
    cost3 = classif("classif3", out3) + cost2
    return cost1, cost2, cost3


def psmnet_hotpath(params: Dict[str, torch.Tensor], fL: torch.Tensor, fR: torch.Tensor, maxdisp: int,
                   out_hw: Tuple[int, int], align_corners: bool = True, operand_dtype=None):
    """The north-star path: PSMNet.forward from the feature maps on (stackhourglass.py:123-168):
    concat volume -> dres0..classif3 -> 3x (trilinear upsample, softmax, regression).
    Returns [pred3, pred2, pred1], each [B,H,W] (the order of :168)."""
    cost = concat_volume(fL, fR, maxdisp // 4, "psm")
    if isinstance(operand_dtype, tuple) and operand_dtype[1] is not None:
        cost = cost.to(operand_dtype[1]).float()
    cost1, cost2, cost3 = psmnet_aggregate(params, cost, operand_dtype)
    size = [maxdisp, out_hw[0], out_hw[1]]
    preds = [upsample_softargmin(c.squeeze(1), size, align_corners) for c in (cost3, cost2, cost1)]
    return preds


def psmnet_random_params(seed: int = 0, calibrate_on: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Synthetic parameters of the 3-D part of PSMNet with the reference's names and init
    (stackhourglass.py:73-114: conv weights N(0, sqrt(2/(k^3*Cout))), BN weight 1 / bias 0).
    With `calibrate_on` (a cost volume) the BatchNorm running statistics are set, layer by layer,
    to the batch statistics of that volume so that eval-mode activations stay O(1) (SURVEY hard
    part 4); otherwise running stats are (0, 1)."""
    rs = np.random.RandomState(seed)     # numpy's MT19937 stream is identical on every host
    params: Dict[str, torch.Tensor] = {}  # (torch's CPU randn is not: it depends on the SIMD width)

    def conv(name, cin, cout, transposed=False):
        n = 27 * cout
        shape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
        # nn.ConvTranspose3d is not an nn.Conv3d instance: the reference's init loop leaves it at
        # torch's default init; we use the same normal law for both (synthetic weights).
        w = rs.standard_normal(size=shape).astype(np.float32) * np.float32(math.sqrt(2.0 / n))
        params[name + ".weight"] = torch.from_numpy(w)

    def bn(name, c):
        params[name + ".weight"] = torch.ones(c)
        params[name + ".bias"] = torch.zeros(c)
        params[name + ".running_mean"] = torch.zeros(c)
        params[name + ".running_var"] = torch.ones(c)

    def convbn(name, cin, cout, transposed=False):
        conv(name + ".0", cin, cout, transposed); bn(name + ".1", cout)

    convbn("dres0.0", 64, 32); convbn("dres0.2", 32, 32)
    convbn("dres1.0", 32, 32); convbn("dres1.2", 32, 32)
    for h in ("dres2", "dres3", "dres4"):
        convbn(h + ".conv1.0", 32, 64); convbn(h + ".conv2", 64, 64)
        convbn(h + ".conv3.0", 64, 64); convbn(h + ".conv4.0", 64, 64)
        convbn(h + ".conv5", 64, 64, True); convbn(h + ".conv6", 64, 32, True)
    for c in ("classif1", "classif2", "classif3"):
        convbn(c + ".0", 32, 32); conv(c + ".2", 32, 1)
    if calibrate_on is not None:
        calibrate_bn_(params, calibrate_on)
    return params


def calibrate_bn_(params: Dict[str, torch.Tensor], cost: torch.Tensor) -> None:
    """Set every BatchNorm's running_mean/var to the statistics its input conv produces on
    `cost` (what one train-mode pass with momentum 1 would store; biased variance is used, the
    synthetic harness only needs O(1) activations).  Walks the graph in execution order."""
    def stat(prefix, y):
        params[prefix + ".running_mean"] = y.mean(dim=(0, 2, 3, 4))
        params[prefix + ".running_var"] = y.var(dim=(0, 2, 3, 4), unbiased=False) + 1e-3

    def cb(prefix, x, stride=1, transposed=False, residual=None, relu=False):
        w = params[prefix + ".0.weight"]
        raw = conv3d_block(x, w, None, None, stride, transposed)
        stat(prefix + ".1", raw)
        return _convbn(params, prefix, x, stride, transposed, residual, relu)

    def hg(p, x, presqu, postsqu):
        out = cb(p + ".conv1.0", x, stride=2, relu=True)
        pre = cb(p + ".conv2", out, residual=postsqu, relu=True)
        out = cb(p + ".conv3.0", pre, stride=2, relu=True)
        out = cb(p + ".conv4.0", out, relu=True)
        post = cb(p + ".conv5", out, transposed=True, residual=presqu if presqu is not None else pre, relu=True)
        out = cb(p + ".conv6", post, transposed=True)
        return out, pre, post

    c0 = cb("dres0.0", cost, relu=True); c0 = cb("dres0.2", c0, relu=True)
    t = cb("dres1.0", c0, relu=True); cost0 = cb("dres1.2", t, residual=c0)
    out1, pre1, post1 = hg("dres2", cost0, None, None); out1 = crop_add(out1, cost0)
    out2, pre2, post2 = hg("dres3", out1, pre1, post1); out2 = crop_add(out2, cost0)
    out3, pre3, post3 = hg("dres4", out2, pre1, post2); out3 = crop_add(out3, cost0)
    for c, x in (("classif1", out1), ("classif2", out2), ("classif3", out3)):
        cb(c + ".0", x, relu=True)


def psmnet_matcher_params(seed: int = 0, noise: float = 0.15, sharpness: float = 1.0) -> Dict[str, torch.Tensor]:
    """Synthetic parameters that make the PSMNet 3-D stack an actual (crude) stereo matcher, so that
    an end-point error against a known disparity is a meaningful number (the north_star's bf16
    tolerance is stated on mean EPE; with purely random weights the EPE is tens of pixels and
    says nothing).  Construction, on top of He-normal noise scaled by `noise`:
      dres0.0  centre tap computes relu(+-(fL_c - fR_c)) for the first 16 feature channels
               -> the 32 outputs sum to a 16-channel sum of absolute differences (SAD);
      dres0.2  3x3 in-plane box filter of each channel (cost aggregation);
      dres1, hourglasses  reference init (stackhourglass.py:100-114); conv6 and dres1.2 scaled by
               `noise`, so that every layer computes full-magnitude activations whose effect on
               the cost is a perturbation;
      classif*.0 identity centre tap + noise; classif*.2 = -sharpness on every channel's centre tap
               -> cost = -sharpness * aggregated SAD, soft-argmin peaks at the matching disparity.
    BatchNorm is the identity (running stats 0/1, weight 1, bias 0)."""
    params = psmnet_random_params(seed)
    rs = np.random.RandomState(seed + 7919)

    def he(shape, cout):
        return torch.from_numpy(rs.standard_normal(size=shape).astype(np.float32) * np.float32(math.sqrt(2.0 / (27 * cout))))

    w = he((32, 64, 3, 3, 3), 32) * noise
    for c in range(16):
        w[c, c, 1, 1, 1] += 1.0; w[c, 32 + c, 1, 1, 1] -= 1.0
        w[16 + c, c, 1, 1, 1] -= 1.0; w[16 + c, 32 + c, 1, 1, 1] += 1.0
    params["dres0.0.0.weight"] = w
    w = he((32, 32, 3, 3, 3), 32) * noise
    for c in range(32):
        w[c, c, 1, :, :] += 1.0 / 9.0
    params["dres0.2.0.weight"] = w
    params["dres1.2.0.weight"] = params["dres1.2.0.weight"] * noise
    for h in ("dres2", "dres3", "dres4"):
        params[h + ".conv6.0.weight"] = params[h + ".conv6.0.weight"] * noise
    for cname in ("classif1", "classif2", "classif3"):
        w = he((32, 32, 3, 3, 3), 32) * noise
        for c in range(32):
            w[c, c, 1, 1, 1] += 1.0
        params[cname + ".0.0.weight"] = w
        w = he((1, 32, 3, 3, 3), 1) * (noise * 0.05)
        w[0, :, 1, 1, 1] -= sharpness
        params[cname + ".2.weight"] = w
    return params


def synthetic_stereo_features(H: int, W: int, C: int = 32, d_lo: float = 10.0, d_hi: float = 30.0, seed: int = 0):
    """A feature pair with stereo structure: fR is fL resampled by a smooth (linear-ramp) disparity
    field, so that fL[y, x] ~ fR[y, x - disp(x)].  Returns (fL, fR, disp) with disp [1,H,W] in units
    of feature-map pixels."""
    rs = np.random.RandomState(seed)
    fL = torch.from_numpy(rs.standard_normal(size=(1, C, H, W)).astype(np.float32))
    # smooth the features a little along x so that sub-pixel interpolation is meaningful
    fL = F.avg_pool2d(F.pad(fL, (1, 1, 0, 0), mode="replicate"), (1, 3), stride=1) * math.sqrt(3.0)
    xr = torch.arange(W, dtype=torch.float32)
    # disparity as a function of the RIGHT-image column x' ; the left column is x = x' + d(x')
    d_r = d_lo + (d_hi - d_lo) * xr / (W - 1)
    src_x = (xr + d_r).clamp(0, W - 1)                       # fR[x'] = fL[x' + d(x')]
    gx = (src_x / (W - 1) * 2 - 1).view(1, 1, W).expand(1, H, W)
    gy = (torch.arange(H, dtype=torch.float32) / max(H - 1, 1) * 2 - 1).view(1, H, 1).expand(1, H, W)
    fR = F.grid_sample(fL, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="border", align_corners=True)
    # disparity seen from the LEFT image: solve x = x' + d(x') for x' (d is linear in x')
    a = (d_hi - d_lo) / (W - 1)
    xl = torch.arange(W, dtype=torch.float32)
    xprime = (xl - d_lo) / (1 + a)
    disp = (xl - xprime).view(1, 1, W).expand(1, H, W).contiguous()
    return fL, fR, disp


# ------------------------------------------------------------------------------------------
# GC-Net 3-D path: concat volume -> feature3d enc-dec -> soft-argmin of MINUS the cost
# ------------------------------------------------------------------------------------------

GC_LAYERS = [  # name, Cin, Cout, stride, transposed  (gcnet.py:38-61)
    ("l19", 64, 32, 1, False), ("l20", 32, 32, 1, False),
    ("l21", 64, 64, 2, False), ("l22", 64, 64, 1, False), ("l23", 64, 64, 1, False),
    ("l24", 64, 64, 2, False), ("l25", 64, 64, 1, False), ("l26", 64, 64, 1, False),
    ("l27", 64, 64, 2, False), ("l28", 64, 64, 1, False), ("l29", 64, 64, 1, False),
    ("l30", 64, 128, 2, False), ("l31", 128, 128, 1, False), ("l32", 128, 128, 1, False),
    ("l33", 128, 64, 2, True), ("l34", 64, 64, 2, True), ("l35", 64, 64, 2, True), ("l36", 64, 32, 2, True),
    ("l37", 32, 1, 2, True),
]


def _gc_block(params, name, x, stride=1, transposed=False, residual=None, operand_dtype=None):
    """conv3d_bn / deconv3d_bn (util_conv.py:150-179): conv(+bias) -> BatchNorm3d (eval) -> ReLU; a skip,
    when there is one, is added AFTER the activation (myAdd3d at gcnet.py:78,84,90,96).  l37 has neither
    BatchNorm nor activation and is a bare module (`l37.weight`), the others are Sequentials (`l19.0.weight`)."""
    bare = (name + ".weight") in params
    w = params[name + (".weight" if bare else ".0.weight")]
    b = params.get(name + (".bias" if bare else ".0.bias"))
    cout = w.shape[1] if transposed else w.shape[0]
    bn = None if bare else _bn(params, name + ".1")
    scale, shift = fold_bn(cout, bn, b)
    od, sd = operand_dtype if isinstance(operand_dtype, tuple) else (operand_dtype, None)
    if bare:
        sd = None                                     # the single-channel output stays fp32
    return conv3d_block(x, w, scale, shift, stride, transposed, residual, 0 if bare else 2, od, sd)


def gcnet_aggregate(params: Dict[str, torch.Tensor], cost: torch.Tensor, operand_dtype=None) -> torch.Tensor:
    """feature3d.forward up to x37 (gcnet.py:65-101), eval-mode BatchNorm.  cost: [B,64,D,H,W] ->
    [B,1,2D,2H,2W]."""
    od = operand_dtype
    g = lambda n, x, s=1, t=False, r=None: _gc_block(params, n, x, s, t, r, od)
    x21 = g("l21", cost, 2); x24 = g("l24", x21, 2); x27 = g("l27", x24, 2); x30 = g("l30", x27, 2)
    x32 = g("l32", g("l31", x30))
    x29 = g("l29", g("l28", x27))
    x33 = g("l33", x32, 2, True, x29)
    x26 = g("l26", g("l25", x24))
    x34 = g("l34", x33, 2, True, x26)
    x23 = g("l23", g("l22", x21))
    x35 = g("l35", x34, 2, True, x23)
    x20 = g("l20", g("l19", cost))
    x36 = g("l36", x35, 2, True, x20)
    return g("l37", x36, 2, True)


def gcnet_hotpath(params: Dict[str, torch.Tensor], fL: torch.Tensor, fR: torch.Tensor, maxdisp: int, operand_dtype=None):
    """gcnet.forward from the feature maps on (gcnet.py:129-137, feature3d :65-111): GC volume with
    D = maxdisp/2, 3-D enc-dec, softmax over MINUS the cost, regression.  Returns [B,1,2H,2W]."""
    cost = concat_volume(fL, fR, maxdisp // 2, "gc")
    if isinstance(operand_dtype, tuple) and operand_dtype[1] is not None:
        cost = cost.to(operand_dtype[1]).float()
    x37 = gcnet_aggregate(params, cost, operand_dtype)
    return softargmin(x37.squeeze(1), -1.0).unsqueeze(1)


def gcnet_random_params(seed: int = 0, calibrate_on: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Synthetic parameters of GC-Net's feature3d with the reference's names (`l19.0.weight`, `l19.0.bias`,
    `l19.1.running_var`, ..., bare `l37.weight/bias`); He-normal weights, small biases; with `calibrate_on`
    (a GC cost volume) the BatchNorm running statistics are set to the batch statistics layer by layer."""
    rs = np.random.RandomState(seed)
    params: Dict[str, torch.Tensor] = {}
    for name, cin, cout, stride, transposed in GC_LAYERS:
        shape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
        taps = 27.0 / 8.0 if transposed else 27.0          # taps that reach one output voxel on average
        w = rs.standard_normal(size=shape).astype(np.float32) * np.float32(math.sqrt(2.0 / (taps * cin)))
        b = rs.standard_normal(size=(cout,)).astype(np.float32) * np.float32(0.05)
        if name == "l37":
            params[name + ".weight"] = torch.from_numpy(w); params[name + ".bias"] = torch.from_numpy(b)
            continue
        params[name + ".0.weight"] = torch.from_numpy(w); params[name + ".0.bias"] = torch.from_numpy(b)
        params[name + ".1.weight"] = torch.ones(cout); params[name + ".1.bias"] = torch.zeros(cout)
        params[name + ".1.running_mean"] = torch.zeros(cout); params[name + ".1.running_var"] = torch.ones(cout)
    if calibrate_on is not None:
        def cal(name, x, stride=1, transposed=False, residual=None):
            w = params[name + ".0.weight"]; b = params[name + ".0.bias"]
            raw = conv3d_block(x, w, None, b, stride, transposed)
            params[name + ".1.running_mean"] = raw.mean(dim=(0, 2, 3, 4))
            params[name + ".1.running_var"] = raw.var(dim=(0, 2, 3, 4), unbiased=False) + 1e-3
            return _gc_block(params, name, x, stride, transposed, residual)
        c = calibrate_on
        x21 = cal("l21", c, 2); x24 = cal("l24", x21, 2); x27 = cal("l27", x24, 2); x30 = cal("l30", x27, 2)
        x32 = cal("l32", cal("l31", x30))
        x29 = cal("l29", cal("l28", x27)); x33 = cal("l33", x32, 2, True, x29)
        x26 = cal("l26", cal("l25", x24)); x34 = cal("l34", x33, 2, True, x26)
        x23 = cal("l23", cal("l22", x21)); x35 = cal("l35", x34, 2, True, x23)
        x20 = cal("l20", cal("l19", c)); cal("l36", x35, 2, True, x20)
    return params


# ------------------------------------------------------------------------------------------
# training-mode restatement (BatchNorm with batch statistics), differentiable with torch autograd
# ------------------------------------------------------------------------------------------

class _RoundST(torch.autograd.Function):
    """Number-format emulation for the training path: forward rounds to `dtype` (straight-through), backward rounds the
    gradient that flows through this point to `grad_dtype` — the CUDA path stores activations AND activation gradients
    (conv-output gradients, conv-input gradients, skip gradients) in bf16; weight gradients stay fp32."""

    @staticmethod
    def forward(ctx, t, dtype, grad_dtype):
        ctx.grad_dtype = grad_dtype
        return t if dtype is None else t.to(dtype).float()

    @staticmethod
    def backward(ctx, g):
        return (g if ctx.grad_dtype is None else g.to(ctx.grad_dtype).float()), None, None


def _convbn_train(params, prefix, x, stride=1, transposed=False, residual=None, relu=0, operand_dtype=None, eps=1e-5,
                  grad_dtype=None):
    """convbn_3d in TRAIN mode (submodule.py:16-19 under model.train()): conv -> batch-stat BatchNorm -> [+res] -> [relu].
    `operand_dtype=torch.bfloat16` rounds conv operands and the stored block output like the kernels do (straight-through
    for autograd); `grad_dtype` additionally rounds the activation gradients at the same points."""
    def rnd(t):
        return t if operand_dtype is None else _RoundST.apply(t, operand_dtype, grad_dtype)
    w = params[prefix + ".0.weight"]
    if operand_dtype is not None:
        w = _RoundST.apply(w, operand_dtype, None)                # weight gradients are fp32
    y = F.conv_transpose3d(rnd(x), w, None, stride=2, padding=1, output_padding=1) if transposed else \
        F.conv3d(rnd(x), w, None, stride=stride, padding=1)
    y = rnd(y)                                            # the conv kernel stores bf16
    y = F.batch_norm(y, None, None, params[prefix + ".1.weight"], params[prefix + ".1.bias"], True, 0.1, eps)
    if residual is not None:
        y = crop_add(y, residual)
    if relu:
        y = F.relu(y)
    return rnd(y)


def psmnet_hotpath_train(params: Dict[str, torch.Tensor], fL: torch.Tensor, fR: torch.Tensor, maxdisp: int,
                         out_hw: Tuple[int, int], align_corners: bool = True, operand_dtype=None, grad_dtype=None):
    """PSMNet.forward from the feature maps on in TRAIN mode (stackhourglass.py:123-168 under model.train()):
    the graph of psmnet_hotpath with batch-statistics BatchNorm; every tensor in `params` / fL / fR may require grad."""
    od = operand_dtype
    cost = concat_volume(fL, fR, maxdisp // 4, "psm")
    cb = lambda p, x, s=1, t=False, r=None, relu=0: _convbn_train(params, p, x, s, t, r, relu, od, grad_dtype=grad_dtype)
    c0 = cb("dres0.0", cost, relu=1); c0 = cb("dres0.2", c0, relu=1)
    t = cb("dres1.0", c0, relu=1); cost0 = cb("dres1.2", t, r=c0)

    def hg(p, x, presqu, postsqu):
        out = cb(p + ".conv1.0", x, 2, relu=1)
        pre = cb(p + ".conv2", out, r=postsqu, relu=1)
        out = cb(p + ".conv3.0", pre, 2, relu=1)
        out = cb(p + ".conv4.0", out, relu=1)
        post = cb(p + ".conv5", out, t=True, r=presqu if presqu is not None else pre, relu=1)
        return cb(p + ".conv6", post, t=True, r=cost0), pre, post

    out1, pre1, post1 = hg("dres2", cost0, None, None)
    out2, pre2, post2 = hg("dres3", out1, pre1, post1)
    out3, pre3, post3 = hg("dres4", out2, pre1, post2)

    def classif(p, x):
        t = cb(p + ".0", x, relu=1)
        w = params[p + ".2.weight"]
        if od is not None:
            w = _RoundST.apply(w, od, None)
        return F.conv3d(t, w, None, stride=1, padding=1)
    cost1 = classif("classif1", out1); cost2 = classif("classif2", out2) + cost1; cost3 = classif("classif3", out3) + cost2
    size = [maxdisp, out_hw[0], out_hw[1]]
    return [upsample_softargmin(c.squeeze(1), size, align_corners) for c in (cost3, cost2, cost1)]


def gcnet_hotpath_train(params: Dict[str, torch.Tensor], fL: torch.Tensor, fR: torch.Tensor, maxdisp: int, operand_dtype=None,
                        eps: float = 1e-5, grad_dtype=None) -> torch.Tensor:
    """gcnet.forward from the feature maps on in TRAIN mode (gcnet.py:129-137, feature3d :65-111 under model.train()):
    conv+bias -> batch-stat BatchNorm -> ReLU, skip adds after the activation; differentiable with torch autograd."""
    od = operand_dtype

    def rnd(t, gd=grad_dtype):
        return t if od is None else _RoundST.apply(t, od, gd)

    def g(name, x, stride=1, transposed=False, residual=None):
        w, b = params[name + ".0.weight"], params[name + ".0.bias"]
        y = F.conv_transpose3d(rnd(x), rnd(w, None), None, stride=2, padding=1, output_padding=1) if transposed else \
            F.conv3d(rnd(x), rnd(w, None), None, stride=stride, padding=1)
        y = rnd(y) + b.view(1, -1, 1, 1, 1)
        y = F.relu(F.batch_norm(y, None, None, params[name + ".1.weight"], params[name + ".1.bias"], True, 0.1, eps))
        if residual is not None:
            y = crop_add(y, residual)
        return rnd(y)

    cost = concat_volume(fL, fR, maxdisp // 2, "gc")
    x21 = g("l21", cost, 2); x24 = g("l24", x21, 2); x27 = g("l27", x24, 2); x30 = g("l30", x27, 2)
    x32 = g("l32", g("l31", x30))
    x29 = g("l29", g("l28", x27)); x33 = g("l33", x32, 2, True, x29)
    x26 = g("l26", g("l25", x24)); x34 = g("l34", x33, 2, True, x26)
    x23 = g("l23", g("l22", x21)); x35 = g("l35", x34, 2, True, x23)
    x20 = g("l20", g("l19", cost)); x36 = g("l36", x35, 2, True, x20)
    x37 = F.conv_transpose3d(rnd(x36), rnd(params["l37.weight"], None), params["l37.bias"], stride=2, padding=1, output_padding=1)
    return softargmin(x37.squeeze(1), -1.0).unsqueeze(1)


# ------------------------------------------------------------------------------------------
# DispNetC (BASELINE config 1): the caller of op 1, restated functionally for model-level parity
# ------------------------------------------------------------------------------------------

_DISPNETC_ENC = [("conv1", 3, 64, 7, 2), ("conv2", 64, 128, 5, 2), ("redir", 128, 64, 1, 1), ("conv3a", 105, 256, 5, 2),
                 ("conv3b", 256, 256, 3, 1), ("conv4a", 256, 512, 3, 2), ("conv4b", 512, 512, 3, 1), ("conv5a", 512, 512, 3, 2),
                 ("conv5b", 512, 512, 3, 1), ("conv6a", 512, 1024, 3, 2), ("conv6b", 1024, 1024, 3, 1)]
_DISPNETC_DEC = [(5, 1024, 512, 1025), (4, 512, 256, 769), (3, 256, 128, 385), (2, 128, 64, 193), (1, 64, 32, 97)]


def dispnetc_random_params(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic DispNetC parameters under the reference's names, init law of net_init (util_conv.py:32-53) with the
    pr* weights scaled by 0.1 (dispnetcorr.py:63-64); numpy RNG so that every host builds the same tensors."""
    rs = np.random.RandomState(seed)
    p: Dict[str, torch.Tensor] = {}

    def conv(name, cin, cout, k, transposed=False, gain=1.0):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        std = math.sqrt(2.0 / (k * k * cout)) * gain
        p[name + ".weight"] = torch.from_numpy(rs.standard_normal(size=shape).astype(np.float32) * np.float32(std))
        p[name + ".bias"] = torch.from_numpy(rs.standard_normal(size=(cout,)).astype(np.float32) * np.float32(0.01))

    for name, cin, cout, k, s in _DISPNETC_ENC:
        conv(name + ".0", cin, cout, k)
    conv("pr6", 1024, 1, 3, gain=0.1)
    for lvl, cin, cout, icin in _DISPNETC_DEC:
        conv("deconv%d.0" % lvl, cin, cout, 4, transposed=True)
        conv("iconv%d.0" % lvl, icin, cout, 3)
        conv("pr%d" % lvl, cout, 1, 3, gain=0.1)
    return p


def dispnetc_forward(p: Dict[str, torch.Tensor], imL: torch.Tensor, imR: torch.Tensor, mode: str = "train",
                     maxdisparity: int = 192, align_corners: bool = True):
    """dispnetcorr.forward (models/dispnetcorr.py:66-134): returns the 7-level pyramid [pr0 .. pr6]."""
    strides = {n: s for n, _, _, _, s in _DISPNETC_ENC}
    ks = {n: k for n, _, _, k, _ in _DISPNETC_ENC}

    def cr(name, x):
        return F.relu(F.conv2d(x, p[name + ".0.weight"], p[name + ".0.bias"], stride=strides[name], padding=(ks[name] - 1) // 2))

    def up(x):
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=align_corners)

    def cat(*ts):
        h = min(t.shape[2] for t in ts); w = min(t.shape[3] for t in ts)
        return torch.cat([t[:, :, :h, :w] for t in ts], dim=1)

    c1L, c1R = cr("conv1", imL), cr("conv1", imR)
    c2L, c2R = cr("conv2", c1L), cr("conv2", c1R)
    x = torch.cat([corr1d(c2L, c2R, 41, 1, 1), cr("redir", c2L)], dim=1)
    skips = {1: c1L, 2: c2L}
    for lvl in (3, 4, 5, 6):
        x = cr("conv%db" % lvl, cr("conv%da" % lvl, x)); skips[lvl] = x
    out = [F.conv2d(x, p["pr6.weight"], p["pr6.bias"], padding=1)]
    for lvl, _, _, _ in _DISPNETC_DEC:
        d = F.relu(F.conv_transpose2d(x, p["deconv%d.0.weight" % lvl], p["deconv%d.0.bias" % lvl], stride=2, padding=1))
        x = F.relu(F.conv2d(cat(d, up(out[0]), skips[lvl]), p["iconv%d.0.weight" % lvl], p["iconv%d.0.bias" % lvl], padding=1))
        out.insert(0, F.conv2d(x, p["pr%d.weight" % lvl], p["pr%d.bias" % lvl], padding=1))
    out.insert(0, up(out[0])[:, :, :imL.shape[-2], :imL.shape[-1]])
    if mode == "test":
        out[-1] = out[-1].clamp(1e-6, max(maxdisparity, imL.shape[-1]))
    return out


# ------------------------------------------------------------------------------------------
# 2-D trunks: PSMNet feature_extraction (submodule.py:65-140), GC-Net feature2d (gcnet.py:14-29)
# ------------------------------------------------------------------------------------------

def _round(t, dtype):
    return t if dtype is None else t.to(dtype).float()


def conv2d_block(x, weight, scale=None, shift=None, stride=1, dilation=1, residual=None, relu=False,
                 operand_dtype=None, storage_dtype=None, padding=None):
    """Conv2d(k, stride, padding = dilation*(k//2), dilation, bias folded into `shift`) -> per-channel affine (eval-mode
    BatchNorm2d) -> [+ residual] -> [ReLU].  convbn (submodule.py:10-13; note its `pad` argument is ignored in favour of
    `dilation`), BasicBlock (:21-42: NO ReLU after the add), util_conv.conv2d_bn (:119-132) / BasicBlock (:180-208: ReLU
    after the add).  `operand_dtype` / `storage_dtype` emulate the tensor-core kernel's number format as in conv3d_block."""
    k = weight.shape[-1]
    if padding is None:
        padding = dilation * (k // 2) if k > 1 else 0
    y = F.conv2d(_round(x, operand_dtype), _round(weight, operand_dtype), None, stride=stride, padding=padding, dilation=dilation)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return _round(y, storage_dtype)


def _fold2d(params, conv_key, bn_prefix, eps=1e-5):
    w = params[conv_key + ".weight"]
    b = params.get(conv_key + ".bias")
    bn = None if bn_prefix is None else {k: params[bn_prefix + "." + k] for k in ("weight", "bias", "running_mean", "running_var")}
    scale, shift = fold_bn(w.shape[0], bn, b, eps)
    return w, scale, shift


def psmnet_feature_extraction(params: Dict[str, torch.Tensor], x: torch.Tensor, align_corners: bool = True,
                              operand_dtype=None, taps: Optional[dict] = None) -> torch.Tensor:
    """feature_extraction.forward (submodule.py:119-140), eval-mode BatchNorm.  `params` uses the reference's names without
    the `feature_extraction.` prefix (firstconv.0.0.weight ...).  operand_dtype = (operand, storage) emulates the CUDA trunk:
    bf16 conv operands and bf16 activation storage, fp32 accumulation; the SPP pooling / 1x1 convs / bilinear upsampling run
    in fp32 on the stored bf16 skip tensor and are stored as bf16; the final 32-channel map is fp32.  `taps` (dict) receives
    intermediate tensors for layer-by-layer debugging."""
    od, sd = operand_dtype if isinstance(operand_dtype, tuple) else (operand_dtype, None)

    def cb(prefix, t, stride=1, dil=1, residual=None, relu=False, storage=sd):
        w, sc, sh = _fold2d(params, prefix + ".0", prefix + ".1")
        return conv2d_block(t, w, sc, sh, stride, dil, residual, relu, od, storage)

    def layer(name, t, blocks, stride, dil):
        for i in range(blocks):
            p = "%s.%d" % (name, i)
            s = stride if i == 0 else 1
            h = cb(p + ".conv1.0", t, s, dil, relu=True)
            res = t
            if (p + ".downsample.0.weight") in params:
                w, sc, sh = _fold2d(params, p + ".downsample.0", p + ".downsample.1")
                res = conv2d_block(t, w, sc, sh, s, 1, None, False, od, sd)
            t = cb(p + ".conv2", h, 1, dil, residual=res)             # BasicBlock: no ReLU after the add (submodule.py:40)
        return t

    # the first convolution reads the fp32 image with fp32 weights (CUDA-core kernel; K = 27 is no tensor-core shape)
    w, sc, sh = _fold2d(params, "firstconv.0.0", "firstconv.0.1")
    t = conv2d_block(x, w, sc, sh, 2, 1, None, True, None, sd)
    t = cb("firstconv.2", t, relu=True)
    t = cb("firstconv.4", t, relu=True)
    if taps is not None: taps["firstconv"] = t
    t = layer("layer1", t, 3, 1, 1)
    if taps is not None: taps["layer1"] = t
    raw = layer("layer2", t, 16, 2, 1)
    if taps is not None: taps["raw"] = raw
    t = layer("layer3", raw, 3, 1, 1)
    if taps is not None: taps["layer3"] = t
    skip = layer("layer4", t, 3, 1, 2)
    if taps is not None: taps["skip"] = skip
    hw = skip.shape[2:]
    br = []
    for i, k in ((1, 64), (2, 32), (3, 16), (4, 8)):
        p = F.avg_pool2d(skip, (k, k), stride=(k, k))
        w, sc, sh = _fold2d(params, "branch%d.1.0" % i, "branch%d.1.1" % i)
        # convbn ignores its `pad` argument and pads by `dilation` = 1 even for this 1x1 convolution (submodule.py:12,84-98):
        # the pooled map grows by a ring whose value is relu(BatchNorm(0)); fp32 on the CUDA cores (a handful of pixels)
        p = conv2d_block(p, w, sc, sh, 1, 1, None, True, padding=1)
        br.append(_round(F.interpolate(p, hw, mode="bilinear", align_corners=align_corners), sd))
        if taps is not None: taps["branch%d" % i] = br[-1]
    feat = torch.cat((raw, skip, br[3], br[2], br[1], br[0]), 1)
    t = cb("lastconv.0", feat, relu=True)
    if taps is not None: taps["lastconv0"] = t
    return conv2d_block(t, params["lastconv.2.weight"], None, None, 1, 1, None, False, od, None)


def gcnet_feature2d(params: Dict[str, torch.Tensor], x: torch.Tensor, operand_dtype=None) -> torch.Tensor:
    """feature2d.forward (gcnet.py:25-29): conv 5x5 s2 + BN + ReLU, 8 BasicBlocks (util_conv.py:180-208, ReLU after the add),
    bare conv 3x3 with bias.  `params` without the `layer2d.` prefix."""
    od, sd = operand_dtype if isinstance(operand_dtype, tuple) else (operand_dtype, None)
    w, sc, sh = _fold2d(params, "conv1.0", "conv1.1")
    t = conv2d_block(x, w, sc, sh, 2, 1, None, True, None, sd)
    for i in range(8):
        p = "block1.%d" % i
        w, sc, sh = _fold2d(params, p + ".conv1", p + ".bn1")
        h = conv2d_block(t, w, sc, sh, 1, 1, None, True, od, sd)
        w, sc, sh = _fold2d(params, p + ".conv2", p + ".bn2")
        t = conv2d_block(h, w, sc, sh, 1, 1, t, True, od, sd)
    w, sc, sh = _fold2d(params, "conv2", None)
    return conv2d_block(t, w, sc, sh, 1, 1, None, False, od, None)


def trunk_random_params(shapes: Dict[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic parameters for a 2-D trunk given {name: shape} (the reference module's state_dict layout): He-normal conv
    weights (fan-out, as stackhourglass.py:100-114 / util_conv.net_init), small biases, BatchNorm weight ~U(0.8,1.2), bias
    ~N(0,0.05), running statistics (0, 1) — calibrate with `calibrate_trunk_bn_`.  numpy RNG: identical on every host."""
    rs = np.random.RandomState(seed)
    p: Dict[str, torch.Tensor] = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            continue
        if len(shp) == 4:
            std = math.sqrt(2.0 / (shp[2] * shp[3] * shp[0]))
            p[name] = torch.from_numpy(rs.standard_normal(size=shp).astype(np.float32) * np.float32(std))
        elif name.endswith("running_mean"):
            p[name] = torch.zeros(shp)
        elif name.endswith("running_var"):
            p[name] = torch.ones(shp)
        elif name.endswith(".bias") and (name[: -len(".bias")] + ".running_mean") not in shapes:
            p[name] = torch.from_numpy(rs.standard_normal(size=shp).astype(np.float32) * np.float32(0.02))     # conv bias
        elif name.endswith(".weight"):
            p[name] = torch.from_numpy(rs.uniform(0.8, 1.2, size=shp).astype(np.float32))
        else:
            p[name] = torch.from_numpy(rs.standard_normal(size=shp).astype(np.float32) * np.float32(0.05))
    return p
