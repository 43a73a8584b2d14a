"""GPU (B200) parity tests: every CUDA op, called through the C-ABI, against the CPU oracle on the
same seeded inputs and against the golden vectors frozen from the reference's own code.

Tolerances (BASELINE.json north_star): bit-exact for concat-volume construction and warp
indexing; <= 1e-4 relative for the fp32 correlation, regression and warp outputs."""
import numpy as np
import pytest
import torch

import oracle.ops as O
from conftest import load_golden, rel_err, rel_err_elem

pytestmark = pytest.mark.gpu
REL = 1e-4


def dev(t):
    return t.cuda()


# ---------------------------------------------------------------- op 1: corr1d
@pytest.mark.parametrize("name", ["corr1d_dispnetc", "corr1d_iresnet2", "corr1d_d_gt_w"])
def test_corr1d_golden(name):
    from dsmnet_b200.corr1d import Corr1d
    g = load_golden(name)
    fL = dev(g["fL"]).requires_grad_(); fR = dev(g["fR"]).requires_grad_()
    out = Corr1d(g["kernel_size"], g["stride"], g["D"])(fL, fR)
    assert rel_err(out, g["out"]) < REL
    assert rel_err_elem(out, g["out"]) < REL          # element-wise, floor at 1 % of the tensor's scale
    out.backward(dev(g["gout"]))
    assert rel_err(fL.grad, g["gL"]) < REL and rel_err(fR.grad, g["gR"]) < REL


@pytest.mark.parametrize("B,C,H,W,D,s", [
    (1, 128, 96, 312, 41, 1),     # DispNetC @384x1248 (dispnetcorr.py:27,77)  — BASELINE config 1
    (1, 128, 135, 240, 81, 1),    # iResNet stage 1 @540x960 (iresnet.py:34,107)
    (1, 64, 270, 480, 41, 2),     # iResNet refinement (iresnet.py:69,175), stride 2
    (2, 20, 7, 45, 13, 1),        # ragged: C, W not multiples of anything
    (1, 6, 3, 9, 5, 2),           # tiny, W%4 != 0 -> scalar store path
])
def test_corr1d_vs_oracle(B, C, H, W, D, s):
    from dsmnet_b200.corr1d import corr1d
    torch.manual_seed(0)
    fL = torch.relu(torch.randn(B, C, H, W)); fR = torch.relu(torch.randn(B, C, H, W))
    ref = O.corr1d(fL, fR, D, s)
    a = dev(fL).requires_grad_(); b = dev(fR).requires_grad_()
    out = corr1d(a, b, D, s)
    assert rel_err(out, ref) < REL
    assert rel_err_elem(out, ref) < REL          # element-wise, floor at 1 % of the tensor's scale
    g = torch.randn(B, D, H, W)
    gL, gR = O.corr1d_grads(g, fL, fR, s)
    out.backward(dev(g))
    assert rel_err(a.grad, gL) < REL and rel_err(b.grad, gR) < REL


def test_dispnetc_model_golden():
    """BASELINE config 1 at model level: the drop-in dispnetcorr (Corr1d on the sm_100a kernel, stock 2-D convs) vs the
    pyramid the reference's own dispnetcorr produced on CPU (fixture), same parameters through load_state_dict."""
    from dsmnet_b200.dispnetcorr import dispnetcorr
    g = load_golden("dispnetc_forward")
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False     # fp32 reference
    try:
        m = dispnetcorr(192)
        missing = m.load_state_dict(O.dispnetc_random_params(seed=g["seed"]), strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            scales, outs = m(dev(g["imL"]), dev(g["imR"]), "test")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert scales == list(range(7)) and len(outs) == 7
    for i, o in enumerate(outs):
        assert o.shape == g["out%d" % i].shape and rel_err(o, g["out%d" % i]) < 1e-3


def test_iresnet_model_golden():
    """BASELINE config 4 at model level: the drop-in iresnet (both Corr1d call sites and the imwrap feature warp on the
    sm_100a kernels, stock 2-D convs) vs the 10 outputs of the reference's own iresnet (CPU fixture)."""
    from dsmnet_b200.iresnet import iresnet
    from helpers import random_state_dict
    g = load_golden("iresnet_forward")
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = iresnet(192)
        m.load_state_dict(random_state_dict(m, g["seed"]), strict=True)
        m = m.cuda().eval()
        with torch.no_grad():
            torch.manual_seed(g["rng_seed"])
            scales, outs = m(dev(g["imL"]), dev(g["imR"]), "test")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert scales == [int(x) for x in g["scales"]] and len(outs) == 10
    for i, o in enumerate(outs):
        assert o.shape == g["out%d" % i].shape and rel_err(o, g["out%d" % i]) < 2e-3


def test_selfsup_pyramid_loss_golden():
    """the reference's own `depthmono-mask` pyramid loss (fixture: loss value and gradients of all 14 disparity maps) vs
    dsmnet_b200.selfsup with the 28 warps on the sm_100a imwrap kernels (fwd + bwd)"""
    import os
    from conftest import GOLDEN
    from dsmnet_b200.selfsup import losses_pyramid1
    z = np.load(os.path.join(GOLDEN, "selfsup_loss.npz"))
    g = {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}
    ne = g["nedge"]; batch = dev(g["batch"]); b1 = torch.flip(batch, dims=[3])
    crop = lambda t: t[:, :, ne:-ne, ne:-ne].contiguous()
    d = [dev(g["disp%d" % l]).requires_grad_() for l in range(7)]
    d1 = [dev(g["disp1_%d" % l]).requires_grad_() for l in range(7)]
    torch.manual_seed(g["seed"])                                    # the warps draw delt from the global CPU RNG (imwrap.py:70)
    loss = losses_pyramid1(batch[:, 3:6].contiguous(), crop(batch[:, :3]), d, list(range(7)), (ne, ne),
                           b1[:, :3].contiguous(), crop(b1[:, 3:6]), d1, (ne, ne), g["weight_levels"].tolist(), True)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    for l in range(7):
        assert rel_err(d[l].grad, g["g%d" % l]) < 2e-3 and rel_err(d1[l].grad, g["g1_%d" % l]) < 2e-3


def test_corr1d_linearity_full_size():
    """size-independent property at the BASELINE size: corr(a*fL, fR) = a*corr(fL, fR); corr(fL, fR1+fR2) additive."""
    from dsmnet_b200.corr1d import corr1d
    torch.manual_seed(1)
    fL = torch.randn(1, 128, 96, 312, device="cuda"); r1 = torch.randn_like(fL); r2 = torch.randn_like(fL)
    c1 = corr1d(fL, r1, 41); c2 = corr1d(fL, r2, 41); c12 = corr1d(fL, r1 + r2, 41)
    assert rel_err(c12, c1 + c2) < REL
    assert rel_err(corr1d(2.0 * fL, r1, 41), 2.0 * c1) < 1e-6
    assert float(c1[:, 5, :, :5].abs().max()) == 0.0          # x < d is exactly zero


# ---------------------------------------------------------------- op 2: concat volume
@pytest.mark.parametrize("name,mode,dkey", [("volume_psm", "psm", None), ("volume_gc", "gc", "D"), ("volume_gc_right", "gc_right", "D")])
def test_volume_golden_bit_exact(name, mode, dkey):
    from dsmnet_b200.cost_volume import concat_volume
    g = load_golden(name)
    D = g["maxdisp"] // 4 if dkey is None else g[dkey]
    out = concat_volume(dev(g["fL"]), dev(g["fR"]), D, mode)
    assert torch.equal(out.cpu(), g["out"])


@pytest.mark.parametrize("mode", ["psm", "gc", "gc_right"])
@pytest.mark.parametrize("shape", [(1, 32, 12, 40, 9), (2, 8, 5, 13, 20), (1, 32, 96, 312, 48)])
def test_volume_vs_oracle(mode, shape):
    from dsmnet_b200.cost_volume import concat_volume
    B, C, H, W, D = shape
    torch.manual_seed(2)
    fL = torch.randn(B, C, H, W); fR = torch.randn(B, C, H, W)
    ref = O.concat_volume(fL, fR, D, mode)
    a = dev(fL).requires_grad_(); b = dev(fR).requires_grad_()
    out = concat_volume(a, b, D, mode)
    assert torch.equal(out.detach().cpu(), ref)                       # bit-exact (pure copies)
    if B * C * H * W * D < 5e6:
        g = torch.randn_like(ref)
        gL, gR = O.concat_volume_grads(g, D, mode)
        out.backward(dev(g))
        assert rel_err(a.grad, gL) < 1e-5 and rel_err(b.grad, gR) < 1e-5
    if C % 8 == 0:
        vol = concat_volume(dev(fL), dev(fR), D, mode, padded_bf16=True)
        v6 = vol.view6().float().cpu()                                # [B,D+2,H+2,W+2,2C]
        inner = v6[:, 1:-1, 1:-1, 1:-1, :].permute(0, 4, 1, 2, 3)
        assert torch.equal(inner, ref.to(torch.bfloat16).float())     # RNE of the same values
        rim = v6.clone(); rim[:, 1:-1, 1:-1, 1:-1, :] = 0
        assert float(rim.abs().max()) == 0.0                          # zero rim written
        assert torch.equal(vol.to_ncdhw().cpu(), ref.to(torch.bfloat16).float())


@pytest.mark.parametrize("mode", ["psm", "gc", "gc_right"])
@pytest.mark.parametrize("shape", [(1, 32, 12, 40, 9), (2, 8, 5, 13, 20), (1, 32, 16, 312, 48), (1, 16, 6, 500, 30)])
def test_volume_padded_backward_vs_oracle(mode, shape):
    """differentiable padded-bf16 volume: backward kernel (sum over d of the padded NDHWC bf16 gradient) against the
    oracle's NCDHW gradient of the same bf16-rounded values"""
    from dsmnet_b200.cost_volume import concat_volume_padded
    from dsmnet_b200.volume_layout import PaddedVolume
    B, C, H, W, D = shape
    torch.manual_seed(6)
    fL = torch.randn(B, C, H, W); fR = torch.randn(B, C, H, W)
    a = dev(fL).requires_grad_(); b = dev(fR).requires_grad_()
    vol = concat_volume_padded(a, b, D, mode)
    ref = O.concat_volume(fL, fR, D, mode)
    assert torch.equal(PaddedVolume(vol.data.detach(), *vol.shape5).to_ncdhw().cpu(), ref.to(torch.bfloat16).float())
    g = torch.randn_like(ref).to(torch.bfloat16).float()               # bf16-representable gradient, NCDHW
    gv = PaddedVolume.from_ncdhw(dev(g))
    vol.data.backward(gv.data)
    gL, gR = O.concat_volume_grads(g, D, mode)
    assert rel_err(a.grad, gL) < 1e-5 and rel_err(b.grad, gR) < 1e-5


def test_pack_unpack_roundtrip():
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(3)
    x = torch.randn(2, 32, 5, 7, 11)
    v = PaddedVolume.from_ncdhw(dev(x))
    assert torch.equal(v.to_ncdhw().cpu(), x.to(torch.bfloat16).float())
    v6 = v.view6().float().cpu(); v6[:, 1:-1, 1:-1, 1:-1, :] = 0
    assert float(v6.abs().max()) == 0.0


# ---------------------------------------------------------------- op 4: soft-argmin
def test_heads_golden():
    from dsmnet_b200.softargmin import softargmin, upsample_softargmin, disparityregression
    g = load_golden("head_psm")
    assert rel_err(upsample_softargmin(dev(g["cost_lr"]), g["size"], True), g["pred"]) < REL
    assert rel_err(softargmin(dev(g["upsampled"]).squeeze(1), 1.0), g["pred"]) < REL
    assert rel_err(disparityregression(16)(dev(g["prob"])), g["pred"]) < REL
    g = load_golden("head_gc")
    assert rel_err(softargmin(dev(g["x37"]).squeeze(1), -1.0).unsqueeze(1), g["pred"]) < REL


@pytest.mark.parametrize("B,D,H,W,sign", [(1, 192, 64, 128, 1.0), (2, 96, 17, 23, -1.0), (1, 5, 3, 7, 1.0)])
def test_softargmin_vs_oracle(B, D, H, W, sign):
    from dsmnet_b200.softargmin import softargmin
    torch.manual_seed(4)
    cost = torch.randn(B, D, H, W) * 2
    c = cost.clone().requires_grad_()
    ref = O.softargmin(c, sign)
    g = torch.randn_like(ref)
    ref.backward(g)
    x = dev(cost).requires_grad_()
    out = softargmin(x, sign)
    assert rel_err(out, ref) < REL
    assert rel_err_elem(out, ref) < REL          # element-wise, floor at 1 % of the tensor's scale
    out.backward(dev(g))
    assert rel_err(x.grad, c.grad) < REL


@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("lr,size", [((1, 12, 24, 78), (48, 96, 312)), ((2, 5, 7, 9), (17, 26, 35)), ((1, 48, 96, 312), (192, 384, 1248))])
def test_upsample_softargmin_vs_oracle(lr, size, ac):
    from dsmnet_b200.softargmin import upsample_softargmin
    torch.manual_seed(5)
    cost = torch.randn(*lr) * 2
    out = upsample_softargmin(dev(cost), size, ac)
    if lr[1] == 48:      # BASELINE size: check a strip against the oracle (full-size CPU upsample is slow)
        ref = O.upsample_softargmin(cost, size, ac)[:, 100:140]
        assert rel_err(out[:, 100:140], ref) < REL
    else:
        assert rel_err(out, O.upsample_softargmin(cost, size, ac)) < REL


@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("lr,size", [((1, 12, 24, 78), (48, 96, 312)), ((2, 5, 7, 9), (17, 26, 35)), ((2, 6, 9, 70), (24, 36, 280))])
def test_upsample_softargmin_backward_vs_oracle(lr, size, ac):
    """fused head backward (dsm_upsample_softargmin_bwd) against autograd through the oracle's
    F.interpolate(trilinear) -> softmax -> regression (stackhourglass.py:152-166), fp64 on the CPU"""
    from dsmnet_b200.softargmin import upsample_softargmin
    torch.manual_seed(11)
    cost = torch.randn(*lr) * 2
    g = torch.randn(lr[0], size[1], size[2])
    c64 = cost.double().requires_grad_(True)
    ref = O.upsample_softargmin(c64, size, ac)
    ref.backward(g.double())
    x = dev(cost).requires_grad_(True)
    out = upsample_softargmin(x, size, ac)
    assert out.grad_fn is not None
    assert rel_err(out, ref.float()) < REL
    out.backward(dev(g))
    assert rel_err(x.grad, c64.grad.float()) < REL


def test_upsample_softargmin_backward_full_size_properties():
    """BASELINE size (48x96x312 -> 192x384x1248): the gradient of sum(disp) w.r.t. a constant shift of the cost is 0
    (softmax invariance), and a directional derivative matches a finite difference of the forward kernel"""
    from dsmnet_b200.softargmin import upsample_softargmin
    torch.manual_seed(12)
    cost = (torch.randn(1, 48, 96, 312) * 2).cuda()
    x = cost.clone().requires_grad_(True)
    w = torch.randn(1, 384, 1248, device="cuda")
    (upsample_softargmin(x, (192, 384, 1248), True) * w).sum().backward()
    gsum = x.grad.sum().item()
    assert abs(gsum) < 1e-3 * x.grad.abs().sum().item()
    v = torch.randn_like(cost)
    eps = 1e-2
    with torch.no_grad():
        fp = (upsample_softargmin(cost + eps * v, (192, 384, 1248), True).double() * w).sum()
        fm = (upsample_softargmin(cost - eps * v, (192, 384, 1248), True).double() * w).sum()
    fd = ((fp - fm) / (2 * eps)).item()
    an = (x.grad.double() * v).sum().item()
    assert abs(fd - an) < 2e-3 * max(abs(fd), abs(an), 1.0), (fd, an)


def test_softargmin_properties_full_size():
    """at the BASELINE head size (192 x 384 x 1248): shift invariance and one-hot limit."""
    from dsmnet_b200.softargmin import softargmin
    torch.manual_seed(6)
    cost = torch.randn(1, 192, 384, 1248, device="cuda")
    d = softargmin(cost, 1.0)
    assert rel_err(softargmin(cost + 3.0, 1.0), d) < REL             # softmax shift invariance
    assert float(d.min()) >= 0.0 and float(d.max()) <= 191.0
    idx = torch.randint(0, 192, (1, 1, 384, 1248), device="cuda")
    onehot = torch.full_like(cost, -80.0).scatter_(1, idx, 80.0)
    assert float((softargmin(onehot, 1.0) - idx[:, 0].float()).abs().max()) < 1e-3


# ---------------------------------------------------------------- op 5: imwrap
IMWRAP = ["imwrap_plain", "imwrap_fliplr", "imwrap_lefttop", "imwrap_scale2", "imwrap_intdisp", "imwrap_oob"]


def _indices(disp, h0, w0, LeftTop, scale, fliplr):
    from dsmnet_b200 import _lib
    from dsmnet_b200.imwrap import grid_vectors
    B, _, h, w = disp.shape
    row, col = grid_vectors(h0, w0, h, w, LeftTop, scale)
    d = disp.cuda().contiguous(); r = row.cuda(); c = col.cuda()
    x0 = torch.empty(B, h, w, dtype=torch.int32, device="cuda"); y0 = torch.empty_like(x0)
    _lib.check(_lib.lib().dsm_warp_indices(d.data_ptr(), r.data_ptr(), c.data_ptr(), int(fliplr), x0.data_ptr(), y0.data_ptr(),
                                           B, h0, w0, h, w, _lib.stream_ptr()), "dsm_warp_indices")
    return x0.cpu().numpy(), y0.cpu().numpy(), row, col


@pytest.mark.parametrize("name", IMWRAP)
def test_imwrap_golden(name):
    from dsmnet_b200.imwrap import imwrap_BCHW
    g = load_golden(name)
    src = dev(g["src"]).requires_grad_(); disp = dev(g["disp"]).requires_grad_()
    out = imwrap_BCHW(src, disp, g["fliplr"], list(g["LeftTop"]), g["scale_factor"], delt=g["delt"])
    assert rel_err(out, g["out"]) < REL
    assert rel_err_elem(out, g["out"]) < REL          # element-wise, floor at 1 % of the tensor's scale
    assert torch.equal(out.detach().cpu() != 0, g["out"] != 0)        # validity mask (loss.py:156,199)
    out.backward(dev(g["gout"]))
    assert rel_err(src.grad, g["gsrc"]) < REL
    assert rel_err(disp.grad, g["gdisp"]) < 2e-4
    # bit-exact sampling indices vs the independent closed form
    _, _, h0, w0 = g["src"].shape
    x0, y0, row, col = _indices(g["disp"], h0, w0, tuple(g["LeftTop"]), g["scale_factor"], g["fliplr"])
    _, rx0, ry0 = O.imwrap_closed_form(g["src"].numpy(), g["disp"].numpy(), row.numpy(), col.numpy(), g["fliplr"], g["delt"])
    assert np.array_equal(x0, rx0) and np.array_equal(y0, ry0)


def test_imwrap_iresnet_size():
    """iResNet feature-constancy warp (iresnet.py:169): (1,32,540,960) by (1,1,540,960)."""
    from dsmnet_b200.imwrap import imwrap_BCHW
    torch.manual_seed(7)
    src = torch.rand(1, 32, 540, 960); disp = torch.rand(1, 1, 540, 960) * 96
    disp[:, :, ::3] = disp[:, :, ::3].round()                      # integer disparities on a third of the rows
    ref = O.imwrap(src, disp, delt=5e-5)
    out = imwrap_BCHW(dev(src), dev(disp), delt=5e-5)
    assert rel_err(out, ref) < REL
    assert rel_err_elem(out, ref) < REL          # element-wise, floor at 1 % of the tensor's scale
    x0, y0, row, col = _indices(disp, 540, 960, (0, 0), 1, False)
    _, rx0, ry0 = O.imwrap_closed_form(src[:, :1].numpy(), disp.numpy(), row.numpy(), col.numpy(), False, 5e-5)
    assert np.array_equal(x0, rx0) and np.array_equal(y0, ry0)
    # zero disparity is the identity up to the reference's own normalise/un-normalise round trip
    # (the fp32 sampling coordinate misses the integer by <= 2e-6*W)
    ident = imwrap_BCHW(dev(src), torch.zeros(1, 1, 540, 960, device="cuda"), delt=0.0)
    assert float((ident.cpu() - src).abs().max()) < 3e-4
    assert float((ident.cpu() - O.imwrap(src, torch.zeros(1, 1, 540, 960), delt=0.0)).abs().max()) < 1e-5


def test_imwrap_rng_stream():
    """the drop-in consumes exactly one torch.rand(1) per call, like imwrap.py:70."""
    from dsmnet_b200.imwrap import imwrap_BCHW
    src = torch.rand(1, 1, 8, 8, device="cuda"); disp = torch.zeros(1, 1, 8, 8, device="cuda")
    torch.manual_seed(11); imwrap_BCHW(src, disp); after = torch.rand(1)
    torch.manual_seed(11); torch.rand(1); expect = torch.rand(1)
    assert torch.equal(after, expect)


# ---- the self-supervised loss kernels: batched multi-level warp, fused SSIM (losses/loss.py:449-452, losses/SSIM.py:24-42) ----

def test_batched_warp_equals_the_single_warps():
    """all the warps of a training step in ONE launch: forward bit-identical to consecutive imwrap_BCHW calls (same RNG
    draws in the same order), backward equal up to the order of the float atomics"""
    from dsmnet_b200.imwrap import imwrap_BCHW, imwrap_batched
    torch.manual_seed(11)
    src = torch.rand(2, 3, 48 + 16, 96 + 16, device="cuda")
    jobs = []
    for lvl, sf in ((0, 1), (1, 2), (2, 4)):
        h, w = 48 >> lvl, 96 >> lvl
        d = (torch.rand(2, 1, h, w, device="cuda") * 6 / sf).requires_grad_()
        d1 = (torch.rand(2, 1, h, w, device="cuda") * 6 / sf).requires_grad_()
        jobs += [dict(im_src=src, disp=d, fliplr=False, LeftTop=[8, 8], scale_factor=sf),
                 dict(im_src=d1, disp=d, fliplr=True, LeftTop=[0, 0], scale_factor=1),
                 dict(im_src=d, disp=d1, fliplr=True, LeftTop=[0, 0], scale_factor=1)]
    torch.manual_seed(5)
    outs = imwrap_batched(jobs)
    gs = [torch.randn_like(o) for o in outs]
    leaves = []
    for j in jobs:
        for t in (j["im_src"], j["disp"]):
            if t.requires_grad and all(t is not u for u in leaves):
                leaves.append(t)
    gb = torch.autograd.grad(outs, leaves, gs)
    torch.manual_seed(5)
    outs1 = [imwrap_BCHW(j["im_src"], j["disp"], j["fliplr"], j["LeftTop"], j["scale_factor"]) for j in jobs]
    g1 = torch.autograd.grad(outs1, leaves, gs)
    for a, b in zip(outs, outs1):
        assert torch.equal(a, b)
    for a, b in zip(gb, g1):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 37, 70), (1, 3, 256, 640), (1, 1, 64, 96)])
def test_fused_ssim_vs_reference_formula(B, C, H, W):
    """SSIM.py:24-42 in fp64 (five 11x11 Gaussian convolutions) vs the fused kernel, forward and the gradient w.r.t. the
    second image.  Tolerance: the map is a ratio of cancelling second moments; the fp32 reference itself (stock PyTorch on
    this GPU) sits ~1e-4 from fp64, the kernel must be at least as close."""
    import torch.nn.functional as F
    from dsmnet_b200.selfsup import SsimFunction, _gauss_window
    torch.manual_seed(12)
    a = torch.rand(B, C, H, W, device="cuda")
    b = (a + 0.1 * torch.randn(B, C, H, W, device="cuda")).clamp(0, 1).requires_grad_()

    def ref(a_, b_, dtype):
        w = _gauss_window(11, C, a_.float()).to(dtype)
        a_, b_ = a_.to(dtype), b_.to(dtype)
        mu1, mu2 = F.conv2d(a_, w, padding=5), F.conv2d(b_, w, padding=5)
        s1 = F.conv2d(a_ * a_, w, padding=5) - mu1 * mu1
        s2 = F.conv2d(b_ * b_, w, padding=5) - mu2 * mu2
        s12 = F.conv2d(a_ * b_, w, padding=5) - mu1 * mu2
        return ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s1 + s2 + 9e-4))

    s = SsimFunction.apply(a, b)
    g = torch.randn_like(s)
    (gb,) = torch.autograd.grad(s, b, g)
    b64 = b.detach().double().requires_grad_()
    s64 = ref(a.double(), b64, torch.float64)
    (gb64,) = torch.autograd.grad(s64, b64, g.double())
    s32 = ref(a, b.detach(), torch.float32)
    e_mine, e_torch = float((s - s64).abs().max()), float((s32 - s64).abs().max())
    print("ssim %s: max |ours - fp64| %.2e, |torch fp32 - fp64| %.2e; grad rel err %.2e" % ((B, C, H, W), e_mine, e_torch, rel_err(gb, gb64)))
    assert s.shape == (B, 1, H, W)
    assert e_mine <= max(2.0 * e_torch, 2e-4)
    assert rel_err(gb, gb64) < 2e-3


@pytest.mark.parametrize("B,C,H,W,rim", [(1, 32, 96, 312, 1), (2, 64, 7, 13, 2), (3, 8, 5, 33, 0), (1, 128, 11, 70, 2)])
def test_pack_nhwc_bf16_bit_exact(B, C, H, W, rim):
    """dsm_pack_nhwc_bf16 / _pair: NCHW fp32 -> zero-rimmed NHWC bf16, bit-identical to torch's RNE conversion; the rim is written (zero)"""
    from dsmnet_b200.conv3d import pack_features_nhwc, pack_feature_pair_nhwc
    torch.manual_seed(3)
    a = torch.randn(B, C, H, W, device="cuda") * 3
    b = torch.randn(B, C, H, W, device="cuda")
    def ref(x):
        out = torch.zeros(B, H + 2 * rim, W + 2 * rim, C, device="cuda", dtype=torch.bfloat16)
        out[:, rim:rim + H, rim:rim + W, :] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
        return out
    one = pack_features_nhwc(a, rim)
    pa, pb = pack_feature_pair_nhwc(a, b, rim)
    torch.cuda.synchronize()
    assert torch.equal(one.view(torch.int16), ref(a).view(torch.int16))
    assert torch.equal(pa.view(torch.int16), ref(a).view(torch.int16)) and torch.equal(pb.view(torch.int16), ref(b).view(torch.int16))
