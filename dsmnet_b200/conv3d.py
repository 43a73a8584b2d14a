"""3-D convolution blocks on the tcgen05/TMEM implicit-GEMM kernel (``dsm_conv3d_fwd``).

Host-side half of op 3: weight re-packing, eval-mode BatchNorm folding and the fused-layer call.
A ``FusedConv3d`` is what one ``convbn_3d(...)`` (+ReLU, +residual) of the reference becomes
(models/psmnet/submodule.py:16-19, stackhourglass.py:26-41,45-60; GC-Net util_conv.py:150-179):
  y = [relu]( conv(x) * scale + shift [+ residual] )
with x, y, residual in the padded-NDHWC bf16 layout (``PaddedVolume``) and fp32 accumulation, or
fp32 NCDHW output for the single-channel classifier convolutions.
"""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import _lib
from .volume_layout import PaddedVolume

BN_EPS = 1e-5
KWFOLD = __import__("os").environ.get("DSM_KWFOLD", "1") != "0"


def pack_weight(weight: torch.Tensor, transposed: bool) -> torch.Tensor:
    """[Cout,Cin,3,3,3] (Conv3d) or [Cin,Cout,3,3,3] (ConvTranspose3d) fp32 ->
    bf16 [27][CoutP][Cin], tap = (kd*3+kh)*3+kw, CoutP = max(16, Cout) (zero rows appended)."""
    if weight.dim() != 5 or tuple(weight.shape[2:]) != (3, 3, 3):
        raise _lib.DsmError("only 3x3x3 kernels are supported")
    w = weight.detach().float()
    w = w.permute(2, 3, 4, 1, 0) if transposed else w.permute(2, 3, 4, 0, 1)     # [3,3,3,Cout,Cin]
    cout, cin = w.shape[3], w.shape[4]
    w = w.reshape(27, cout, cin)
    coutp = max(16, cout)
    if coutp != cout:
        w = torch.cat([w, w.new_zeros(27, coutp - cout, cin)], dim=1)
    return w.to(torch.bfloat16).contiguous()


def pack_weight_device(weight: torch.Tensor, mode: int) -> torch.Tensor:
    """One-launch repack of a CUDA fp32 weight (dsm_pack_weight): mode 0 Conv3d [Cout,Cin,3,3,3], 1 ConvTranspose3d
    [Cin,Cout,3,3,3], 2 the stride-1 dgrad filter of a Conv3d weight [Cout,Cin,...] -> a conv with Cin outputs."""
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if mode == 0:
        cout, cin = w.shape[0], w.shape[1]
    else:
        cout, cin = w.shape[1], w.shape[0]
    out = torch.empty(27, max(16, cout), cin, device=w.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().dsm_pack_weight(w.data_ptr(), out.data_ptr(), cout, cin, mode, _lib.stream_ptr(w.device)), "dsm_pack_weight")
    return out


def fold_affine(cout: int, bn=None, bias: Optional[torch.Tensor] = None, eps: float = BN_EPS):
    """Eval-mode BatchNorm3d (+ conv bias) -> per-channel (scale, shift), fp32, padded to CoutP.
    scale = gamma/sqrt(var+eps); shift = beta - mean*scale + bias*scale."""
    dev = bias.device if bias is not None else (bn.weight.device if bn is not None else None)
    if bn is None:
        scale = torch.ones(cout, device=dev)
        shift = torch.zeros(cout, device=dev)
    else:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + (bn.eps if hasattr(bn, "eps") else eps))
        shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if bias is not None:
        shift = shift + bias.detach().float() * scale
    coutp = max(16, cout)
    if coutp != cout:
        scale = torch.cat([scale, scale.new_ones(coutp - cout)])
        shift = torch.cat([shift, shift.new_zeros(coutp - cout)])
    return scale.contiguous(), shift.contiguous()


def conv_out_dims(D, H, W, stride, transposed):
    if transposed:
        return 2 * D, 2 * H, 2 * W
    if stride == 2:
        return (D - 1) // 2 + 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1
    return D, H, W


class FusedConv3d:
    """One fused layer with device-resident packed weights."""

    def __init__(self, weight, bn=None, bias=None, stride=1, transposed=False, relu=False, device=None, variant=0, dgrad_flip=False):
        # relu: False/0 none, True/1 after the residual add (PSMNet), 2 before it (GC-Net skip adds)
        # dgrad_flip: `weight` is a Conv3d weight [Cout,Cin,...] and this layer is its stride-1 input-gradient
        # convolution (Cout inputs -> Cin outputs, flipped taps)
        self.transposed, self.stride, self.relu = bool(transposed), int(stride), int(relu)
        if self.transposed or dgrad_flip:
            self.cin, self.cout = weight.shape[0], weight.shape[1]
            if self.transposed:
                self.stride = 2
        else:
            self.cout, self.cin = weight.shape[0], weight.shape[1]
        device = device if device is not None else weight.device
        # single-output-channel stride-1 layers with 32 inputs (PSMNet classif*.2): kw folded into the output columns
        # (dsm_conv3d_fwd_ex variant bit 8; a third of the MMAs); DSM_KWFOLD=0 keeps the plain 32 -> 16 mapping
        self.kwfold = (KWFOLD and self.cout == 1 and self.cin == 32 and not self.transposed and not dgrad_flip and self.stride == 1)
        if self.kwfold:
            variant |= 256
            wk = weight.detach().float().to(device).contiguous()
            self.w = torch.empty(27, 16, 32, device=device, dtype=torch.bfloat16)
            if wk.is_cuda:
                _lib.check(_lib.lib().dsm_pack_weight(wk.data_ptr(), self.w.data_ptr(), 1, 32, 3, _lib.stream_ptr(wk.device)), "dsm_pack_weight")
            else:
                raise _lib.DsmError("FusedConv3d: the packed weights live on a CUDA device")
        elif weight.is_cuda and torch.device(device) == weight.device:
            self.w = pack_weight_device(weight, 2 if dgrad_flip else (1 if self.transposed else 0))
        elif dgrad_flip:
            self.w = pack_weight(weight.detach().flip(2, 3, 4).transpose(0, 1).contiguous(), False).to(device)
        else:
            self.w = pack_weight(weight, self.transposed).to(device)
        self.identity_affine = bn is None and bias is None
        if self.identity_affine:            # the kernels take NULL for scale/shift: nothing to build or copy
            self.scale = self.shift = None
        else:
            scale, shift = fold_affine(self.cout, bn, bias)
            self.scale, self.shift = scale.to(device), shift.to(device)
        self.variant = variant

    def set_affine(self, scale: torch.Tensor, shift: torch.Tensor):
        """Replace the folded per-channel affine (fp32 [Cout] each; padded to CoutP with identity)."""
        dev = self.w.device
        coutp = max(16, self.cout)
        sc = torch.ones(coutp, device=dev); sh = torch.zeros(coutp, device=dev)
        sc[:scale.numel()] = scale.to(dev).float(); sh[:shift.numel()] = shift.to(dev).float()
        self.scale, self.shift, self.identity_affine = sc.contiguous(), sh.contiguous(), False

    def out_dims(self, x: PaddedVolume):
        return conv_out_dims(x.D, x.H, x.W, self.stride, self.transposed)

    def __call__(self, x: PaddedVolume, out: Union[PaddedVolume, torch.Tensor, None] = None,
                 residual: Union[PaddedVolume, torch.Tensor, None] = None):
        """``out``: a PaddedVolume (bf16) or, for Cout==1, an fp32 [B,D,H,W] tensor; allocated when None.
        Its extent may be a crop of the natural output size (myadd_3d semantics)."""
        _lib.require_cuda(x.data, self.w)
        if x.C != self.cin:
            raise _lib.DsmError("FusedConv3d: expected %d input channels, got %d" % (self.cin, x.C))
        nD, nH, nW = self.out_dims(x)
        f32 = self.cout == 1
        if out is None:
            if residual is not None:
                rD, rH, rW = (residual.D, residual.H, residual.W) if isinstance(residual, PaddedVolume) else residual.shape[-3:]
                nD, nH, nW = min(nD, rD), min(nH, rH), min(nW, rW)
            out = torch.empty(x.B, nD, nH, nW, device=x.data.device, dtype=torch.float32) if f32 else \
                PaddedVolume.empty(x.B, self.cout, nD, nH, nW, x.data.device)
        if f32:
            oD, oH, oW = out.shape[-3:]
            optr = out.data_ptr()
            rptr = 0 if residual is None else residual.data_ptr()
            if residual is not None and tuple(residual.shape[-3:]) != (oD, oH, oW):
                raise _lib.DsmError("residual/out extent mismatch")
        else:
            oD, oH, oW = out.D, out.H, out.W
            optr = out.data.data_ptr()
            rptr = 0
            if residual is not None:
                if (residual.D, residual.H, residual.W, residual.C) != (oD, oH, oW, self.cout):
                    raise _lib.DsmError("residual/out extent mismatch")
                rptr = residual.data.data_ptr()
        _lib.check(_lib.lib().dsm_conv3d_fwd_ex(
            x.data.data_ptr(), self.w.data_ptr(),
            0 if self.identity_affine else self.scale.data_ptr(), 0 if self.identity_affine else self.shift.data_ptr(),
            rptr, optr, x.B, self.cin, self.cout, x.D, x.H, x.W, self.stride, int(self.transposed), int(self.relu),
            _lib.DSM_F32 if f32 else _lib.DSM_BF16, oD, oH, oW, self.variant, _lib.stream_ptr(x.data.device)),
            "dsm_conv3d_fwd")
        return out


def pack_features_nhwc(f: torch.Tensor, rim: int = 1) -> torch.Tensor:
    """NCHW fp32 feature map -> bf16 NHWC with a zero rim, [B, H+2r, W+2r, C] (RNE): the operand layout of `conv_from_features`."""
    _lib.require_cuda(f)
    f = f.contiguous().float()
    B, C, H, W = f.shape
    out = torch.empty(B, H + 2 * rim, W + 2 * rim, C, device=f.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().dsm_pack_nhwc_bf16(f.data_ptr(), out.data_ptr(), B, C, H, W, rim, _lib.stream_ptr(f.device)), "dsm_pack_nhwc_bf16")
    return out


def pack_feature_pair_nhwc(fL: torch.Tensor, fR: torch.Tensor, rim: int = 1):
    """both maps of a pair in one launch (dsm_pack_nhwc_bf16_pair)"""
    _lib.require_cuda(fL, fR)
    fL = fL.contiguous().float(); fR = fR.contiguous().float()
    B, C, H, W = fL.shape
    out = torch.empty(2, B, H + 2 * rim, W + 2 * rim, C, device=fL.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().dsm_pack_nhwc_bf16_pair(fL.data_ptr(), fR.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), B, C, H, W, rim,
                                                  _lib.stream_ptr(fL.device)), "dsm_pack_nhwc_bf16_pair")
    return out[0], out[1]


def conv_from_features(layer: "FusedConv3d", featL: torch.Tensor, featR: torch.Tensor, D: int, mode: str,
                       out: Optional[PaddedVolume] = None) -> PaddedVolume:
    """y = layer(concat_volume(fL, fR, D, mode)) without materialising the volume (dsm_conv3d_volume_fwd): `layer` is the
    64 -> 32 stride-1 block that reads the volume (PSMNet dres0.0, GC-Net l19); featL / featR: bf16 NHWC with a zero rim of
    one pixel, [B, H+2, W+2, 32] (`pack_features_nhwc`, or the 2-D trunk's last layer)."""
    _lib.require_cuda(featL, featR, layer.w)
    if layer.cin != 64 or layer.cout != 32 or layer.stride != 1 or layer.transposed:
        raise _lib.DsmError("conv_from_features: the fused volume convolution is the 64 -> 32 stride-1 layer")
    if featL.dtype != torch.bfloat16 or featL.shape != featR.shape or featL.dim() != 4 or featL.shape[3] != 32:
        raise _lib.DsmError("conv_from_features: two zero-rimmed bf16 NHWC [B, H+2, W+2, 32] feature maps expected")
    featL = featL.contiguous(); featR = featR.contiguous()
    B, H, W = featL.shape[0], featL.shape[1] - 2, featL.shape[2] - 2
    if out is None:
        out = PaddedVolume.empty(B, 32, D, H, W, featL.device)
    elif out.shape5 != (B, 32, D, H, W):
        raise _lib.DsmError("conv_from_features: output geometry mismatch")
    _lib.check(_lib.lib().dsm_conv3d_volume_fwd(
        featL.data_ptr(), featR.data_ptr(), layer.w.data_ptr(),
        0 if layer.identity_affine else layer.scale.data_ptr(), 0 if layer.identity_affine else layer.shift.data_ptr(),
        out.data.data_ptr(), B, 32, 32, int(D), H, W, _lib.VOLUME_MODES[mode], int(layer.relu), layer.variant,
        _lib.stream_ptr(featL.device)), "dsm_conv3d_volume_fwd")
    return out


def conv_timeouts() -> int:
    """Pipeline waits that timed out inside conv kernels since load (0 in a healthy run)."""
    return _lib.lib().dsm_debug_conv_timeouts()


# ------------------------------------------------------------------------------------------------
# Training path: a differentiable bare convolution on padded NDHWC bf16 volumes.
#   forward : the same tcgen05 kernels as inference (identity affine, no activation)
#   dgrad   : those kernels again with transformed weights (see include/dsmnet_b200.h, op 3 training)
#   wgrad   : dsm_conv3d_wgrad
# BatchNorm (batch statistics or frozen), ReLU and the (cropped) skip adds of the training graph are the streaming
# kernels of csrc/bnact.cu, wrapped in dsmnet_b200/train3d.py.
# ------------------------------------------------------------------------------------------------

def _dgrad_layer(weight: torch.Tensor, stride: int, transposed: bool, device) -> "FusedConv3d":
    """The layer that maps gy to gx for y = conv(x, weight)."""
    w = weight.detach()
    if transposed:                    # y = conv_transpose(x, w[Cin][Cout]): gx = conv3d(gy, w, stride 2), w read as [out=Cin][in=Cout]
        return FusedConv3d(w, None, None, 2, False, 0, device)
    if stride == 1:                   # gx = conv3d(gy, w'), w'[ci][co][k] = w[co][ci][2-k]
        return FusedConv3d(w, None, None, 1, False, 0, device, dgrad_flip=True)
    # stride 2: gx = conv_transpose3d(gy, w) cropped to x's extent; w[Cout][Cin] is a ConvTranspose weight with in=Cout
    return FusedConv3d(w, None, None, 2, True, 0, device)


_wgrad_ws = {}
_wgrad_ws_retired = []


def _wgrad_workspace(nbytes: int, device) -> torch.Tensor:
    """One growing scratch buffer per device (partials; for stride-2 layers also the parity-gathered fine tensor)."""
    key = str(device)
    ws = _wgrad_ws.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        if ws is not None:
            _wgrad_ws_retired.append(ws)          # a captured CUDA graph may still point into it
        ws = torch.empty((nbytes + 3) // 4, device=device, dtype=torch.float32)
        _wgrad_ws[key] = ws
    return ws


def conv3d_wgrad(anchor: PaddedVolume, partner: PaddedVolume, stride: int, ca_out: int, cb_out: int) -> torch.Tensor:
    """dW[ca_out][cb_out][3][3][3] (fp32) from the padded volumes; see dsm_conv3d_wgrad for anchor/partner."""
    dev = anchor.data.device
    dw = torch.empty(ca_out, cb_out, 3, 3, 3, device=dev, dtype=torch.float32)
    L = _lib.lib()
    ws = _wgrad_workspace(L.dsm_conv3d_wgrad_workspace_bytes_ex(anchor.B, anchor.C, partner.C, anchor.D, anchor.H, anchor.W, stride), dev)
    _lib.check(L.dsm_conv3d_wgrad(
        anchor.data.data_ptr(), partner.data.data_ptr(), dw.data_ptr(), anchor.B, anchor.C, partner.C,
        anchor.D, anchor.H, anchor.W, partner.D, partner.H, partner.W, stride, ca_out, cb_out, 0, 0, 0,
        ws.data_ptr(), ws.numel() * 4, _lib.stream_ptr(dev)), "dsm_conv3d_wgrad")
    return dw


class Conv3dFunction(torch.autograd.Function):
    """ydata = conv(xdata, weight): k=3, pad=1; `geom` = (B, D, H, W, stride, transposed, (Do, Ho, Wo)).
    xdata / ydata are the flat bf16 storages of PaddedVolumes (zero rims)."""

    @staticmethod
    def forward(ctx, xdata, weight, geom):
        B, D, H, W, stride, transposed, odims = geom
        cin = weight.shape[0] if transposed else weight.shape[1]
        cout = weight.shape[1] if transposed else weight.shape[0]
        x = PaddedVolume(xdata, B, cin, D, H, W)
        layer = FusedConv3d(weight, None, None, stride, transposed, 0, xdata.device)
        y = layer(x, PaddedVolume.empty_zero_rim(B, cout, *odims, xdata.device))
        ctx.save_for_backward(xdata, weight)
        ctx.geom = geom
        return y.data

    @staticmethod
    def backward(ctx, gy):
        xdata, weight = ctx.saved_tensors
        B, D, H, W, stride, transposed, odims = ctx.geom
        cin = weight.shape[0] if transposed else weight.shape[1]
        cout = weight.shape[1] if transposed else weight.shape[0]
        dev = xdata.device
        g = PaddedVolume(gy.contiguous().to(torch.bfloat16), B, cout, *odims)
        x = PaddedVolume(xdata, B, cin, D, H, W)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _dgrad_layer(weight, stride, transposed, dev)(g, PaddedVolume.empty_zero_rim(B, cin, D, H, W, dev)).data
        if ctx.needs_input_grad[1]:
            if transposed:
                gw = conv3d_wgrad(x, g, 2, cin, cout)          # anchor = x (coarse), partner = gy
            else:
                gw = conv3d_wgrad(g, x, stride, cout, cin)     # anchor = gy, partner = x
            gw = gw.to(weight.dtype)
        return gx, gw, None


def conv3d_train(x: PaddedVolume, weight: torch.Tensor, stride: int = 1, transposed: bool = False, out_dims=None) -> PaddedVolume:
    """Differentiable y = conv(x, weight) on padded volumes (both channel counts in {32, 64, 128})."""
    cout = weight.shape[1] if transposed else weight.shape[0]
    nat = conv_out_dims(x.D, x.H, x.W, 2 if transposed else stride, transposed)
    odims = tuple(out_dims) if out_dims is not None else nat
    ydata = Conv3dFunction.apply(x.data, weight, (x.B, x.D, x.H, x.W, 2 if transposed else stride, bool(transposed), odims))
    return PaddedVolume(ydata, x.B, cout, *odims)
