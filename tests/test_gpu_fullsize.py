"""GPU (B200): BASELINE.json's full sizes (PSMNet 384x1248, maxdisp 192: volume 64x48x96x312).

The CPU oracle needs ~10 s per pass at this size, so most checks here are size-independent
properties that are exact in bf16: with a one-hot ("delta") filter a convolution is a pure shift /
subsample / zero-stuffing of its input, which exercises every tile boundary, the zero rim, all
eight stride-2 sub-lattices and all eight transposed-conv parity classes over the whole volume.
One full-size pass of the north-star path is compared with the oracle, and the north_star's
tolerance (mean end-point-error delta < 0.01 px) is asserted there."""
import pytest
import torch
import torch.nn.functional as F

import oracle.ops as O

pytestmark = pytest.mark.gpu
D, H, W = 48, 96, 312


def _delta_weight(cin, cout, tap, transposed):
    kd, kh, kw = tap
    w = torch.zeros(cin, cout, 3, 3, 3) if transposed else torch.zeros(cout, cin, 3, 3, 3)
    for c in range(min(cin, cout)):
        w[c, c, kd, kh, kw] = 1.0
    return w


def _run(x, w, stride, transposed):
    from dsmnet_b200.conv3d import FusedConv3d, conv_timeouts
    from dsmnet_b200.volume_layout import PaddedVolume
    y = FusedConv3d(w.cuda(), None, None, stride, transposed, False)(PaddedVolume.from_ncdhw(x))
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    return y.to_ncdhw()


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (0, 1, 2), (2, 0, 1)])
def test_stride1_delta_filter_is_a_shift(tap):
    torch.manual_seed(0)
    x = torch.randn(1, 32, D, H, W, device="cuda").to(torch.bfloat16).float()
    y = _run(x, _delta_weight(32, 32, tap, False), 1, False)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    kd, kh, kw = tap
    assert torch.equal(y, xp[:, :, kd:kd + D, kh:kh + H, kw:kw + W])


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (0, 1, 2), (1, 2, 0)])
def test_stride2_delta_filter_is_a_subsample(tap):
    torch.manual_seed(1)
    x = torch.randn(1, 32, D, H, W, device="cuda").to(torch.bfloat16).float()
    y = _run(x, _delta_weight(32, 64, tap, False), 2, False)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    kd, kh, kw = tap
    ref = xp[:, :, kd:kd + D:2, kh:kh + H:2, kw:kw + W:2]
    assert y.shape == (1, 64, D // 2, H // 2, W // 2)
    assert torch.equal(y[:, :32], ref) and float(y[:, 32:].abs().max()) == 0.0


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (0, 1, 2), (2, 1, 0)])
def test_transposed_delta_filter_is_zero_stuffing(tap):
    torch.manual_seed(2)
    d, h, w = D // 2, H // 2, W // 2
    x = torch.randn(1, 64, d, h, w, device="cuda").to(torch.bfloat16).float()
    wt = _delta_weight(64, 32, tap, True)
    y = _run(x, wt, 2, True)
    ref = F.conv_transpose3d(x[:, :32], wt[:32].cuda(), stride=2, padding=1, output_padding=1)   # exact: one term per output
    assert y.shape == (1, 32, D, H, W)
    assert torch.equal(y, ref)


def test_concat_volume_checksum_full_size():
    """sum over d of the PSM volume = closed-form weighted sums of the inputs (exact in fp64)."""
    from dsmnet_b200.cost_volume import concat_volume
    torch.manual_seed(3)
    fL = torch.randn(1, 32, H, W, device="cuda"); fR = torch.randn(1, 32, H, W, device="cuda")
    vol = concat_volume(fL, fR, D, "psm").double()
    xs = torch.arange(W, device="cuda")
    cnt = torch.clamp(xs + 1, max=D).double()                           # number of d with d <= x
    assert torch.allclose(vol[:, :32].sum(2), fL.double() * cnt, rtol=0, atol=1e-9)
    right = torch.zeros(1, 32, H, W, device="cuda", dtype=torch.float64)
    for d in range(D):
        right[..., d:] += fR.double()[..., : W - d]
    assert torch.allclose(vol[:, 32:].sum(2), right, rtol=0, atol=1e-9)


def _epe(pred, gt, valid):
    return float((pred - gt)[valid].abs().mean())


def test_north_star_path_full_size():
    """One 384x1248 pair through the CUDA path vs the CPU oracle (fp32 reference arithmetic).

    The 3-D stack carries `psmnet_matcher_params` (a crude but real stereo matcher inside the PSMNet
    architecture, oracle/ops.py) and the features have stereo structure with a known disparity, so
    the end-point error is a meaningful number (~1.5 px) and BASELINE.json's tolerance —
    mean EPE delta < 0.01 px for the bf16 tensor-core convolutions — can be asserted as stated.
    EPE is taken where the PSM volume is defined for every candidate disparity (x >= maxdisp)."""
    from dsmnet_b200.psmnet import PSMNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    maxdisp, HI, WI = 192, 384, 1248
    fL, fR, disp4 = O.synthetic_stereo_features(H, W, d_lo=10.0, d_hi=30.0, seed=4)
    gt = F.interpolate(disp4.unsqueeze(1) * 4, size=(HI, WI), mode="bilinear", align_corners=True)[:, 0]
    valid = torch.zeros_like(gt, dtype=torch.bool)
    valid[:, 8:-8, maxdisp + 8:-8] = True
    params = O.psmnet_matcher_params(seed=21)
    ref = O.psmnet_hotpath(params, fL, fR, maxdisp, (HI, WI))
    emu = O.psmnet_hotpath(params, fL, fR, maxdisp, (HI, WI), operand_dtype=(torch.bfloat16, torch.bfloat16))
    m = PSMNetHotPath(maxdisp)
    m.load_state_dict(params, strict=False)
    m = m.cuda().eval()
    with torch.no_grad():
        preds = m(fL.cuda(), fR.cuda(), (HI, WI))
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    for name, mine, r, e in zip(("pred3", "pred2", "pred1"), preds, ref, emu):
        mine = mine.cpu()
        assert mine.shape == (1, HI, WI) and bool(torch.isfinite(mine).all())
        epe_ref, epe_mine = _epe(r, gt, valid), _epe(mine, gt, valid)
        d_ref = float((mine - r)[valid].abs().mean()); d_emu = float((mine - e)[valid].abs().mean())
        print("%s: EPE ref %.4f ours %.4f (delta %.5f px); mean |ours-ref| %.4f, |ours-emu| %.4f px" %
              (name, epe_ref, epe_mine, abs(epe_mine - epe_ref), d_ref, d_emu))
        assert epe_ref < 3.0                           # the synthetic matcher really matches
        assert abs(epe_mine - epe_ref) < 0.01          # north_star: mean EPE delta < 0.01 px
        assert d_ref < 0.05 and d_emu < 0.02           # per-pixel agreement with the fp32 reference / its bf16 emulation


def test_north_star_path_full_size_random_weights():
    """Same pair, purely random (He-init, BN-calibrated) weights: the network is a chaotic amplifier, so the
    gate is relative — the CUDA path must sit closer to the oracle's bf16-operand emulation than that
    emulation sits to the fp32 reference (i.e. the kernels add no error beyond the operand format)."""
    from dsmnet_b200.psmnet import PSMNetHotPath
    maxdisp, HI, WI = 192, 384, 1248
    fL, fR, _ = O.synthetic_stereo_features(H, W, seed=5)
    cost = O.concat_volume(fL, fR, maxdisp // 4, "psm")
    params = O.psmnet_random_params(seed=21, calibrate_on=cost)
    ref = O.psmnet_hotpath(params, fL, fR, maxdisp, (HI, WI))
    emu = O.psmnet_hotpath(params, fL, fR, maxdisp, (HI, WI), operand_dtype=(torch.bfloat16, torch.bfloat16))
    m = PSMNetHotPath(maxdisp)
    m.load_state_dict(params, strict=False)
    m = m.cuda().eval()
    with torch.no_grad():
        preds = m(fL.cuda(), fR.cuda(), (HI, WI))
    for name, mine, r, e in zip(("pred3", "pred2", "pred1"), preds, ref, emu):
        mine = mine.cpu()
        d_emu = float((mine - e).abs().mean()); d_fmt = float((e - r).abs().mean()); d_ref = float((mine - r).abs().mean())
        print("%s: mean |ours-emu| %.4f, |emu-ref| %.4f, |ours-ref| %.4f px" % (name, d_emu, d_fmt, d_ref))
        assert d_emu < d_fmt and d_ref < 2.0 * d_fmt + 1e-3


# ------------------------------------------------------------------------------------------------------------------
# BASELINE config 3 at its full size: GC-Net 256x512 crop, maxdisp 192 -> volume 64 x 96 x 128 x 256 (half resolution),
# l36 output 32 x 96 x 128 x 256, l37 output fp32 192 x 256 x 512, encoder chain down to 6 x 8 x 16.
# ------------------------------------------------------------------------------------------------------------------
GD, GH, GW = 96, 128, 256


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (0, 2, 1)])
def test_gc_fullsize_l19_delta_filter_is_a_shift(tap):
    """64 -> 32 stride-1 plane-sharing kernel over the whole GC-Net volume extent (the 32-bit flat decode, every band/tile edge)"""
    torch.manual_seed(10)
    x = torch.randn(1, 64, GD, GH, GW, device="cuda").to(torch.bfloat16).float()
    y = _run(x, _delta_weight(64, 32, tap, False), 1, False)
    kd, kh, kw = tap
    xp = F.pad(x[:, :32], (1, 1, 1, 1, 1, 1))
    assert torch.equal(y, xp[:, :, kd:kd + GD, kh:kh + GH, kw:kw + GW])


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (1, 0, 2)])
def test_gc_fullsize_stride2_chain_is_a_subsample(tap):
    """l21 -> l24 -> l27 -> l30 geometry: four stride-2 convolutions 96x128x256 -> 6x8x16 (64 -> 64 ... -> 128), delta filters"""
    torch.manual_seed(11)
    x = torch.randn(1, 64, GD, GH, GW, device="cuda").to(torch.bfloat16).float()
    kd, kh, kw = tap
    ref = x
    cur = x
    for cout in (64, 64, 64, 128):
        cin = cur.shape[1]
        cur = _run(cur, _delta_weight(cin, cout, tap, False), 2, False)
        d, h, w = ref.shape[2:]
        rp = F.pad(ref, (1, 1, 1, 1, 1, 1))
        ref = rp[:, :, kd:kd + d:2, kh:kh + h:2, kw:kw + w:2]
        assert cur.shape[2:] == ref.shape[2:]
        assert torch.equal(cur[:, :64], ref[:, :64]) and (cur.shape[1] == 64 or float(cur[:, 64:].abs().max()) == 0.0)
    assert tuple(cur.shape) == (1, 128, 6, 8, 16)


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (2, 0, 1)])
def test_gc_fullsize_l36_transposed_delta_filter(tap):
    """l36: 64 -> 32 transposed (class-sharing kernel), 48x64x128 -> 96x128x256"""
    torch.manual_seed(12)
    x = torch.randn(1, 64, GD // 2, GH // 2, GW // 2, device="cuda").to(torch.bfloat16).float()
    wt = _delta_weight(64, 32, tap, True)
    y = _run(x, wt, 2, True)
    ref = F.conv_transpose3d(x[:, :32], wt[:32].cuda(), stride=2, padding=1, output_padding=1)
    assert y.shape == (1, 32, GD, GH, GW) and torch.equal(y, ref)


@pytest.mark.parametrize("tap", [(1, 1, 1), (0, 0, 0), (2, 2, 2), (0, 1, 2)])
def test_gc_fullsize_l37_fp32_output(tap):
    """l37: 32 -> 1 transposed with the fp32 [192][256][512] output (one-warp-per-quadrant epilogue, 8-byte pair stores)"""
    from dsmnet_b200.conv3d import FusedConv3d, conv_timeouts
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(13)
    x = torch.randn(1, 32, GD, GH, GW, device="cuda").to(torch.bfloat16).float()
    wt = torch.zeros(32, 1, 3, 3, 3)
    kd, kh, kw = tap
    wt[5, 0, kd, kh, kw] = 1.0                                 # picks channel 5: exact in bf16 x fp32
    y = FusedConv3d(wt.cuda(), None, None, 2, True, False)(PaddedVolume.from_ncdhw(x))
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    ref = F.conv_transpose3d(x[:, 5:6], wt[5:6].cuda(), stride=2, padding=1, output_padding=1)[:, 0]
    assert y.shape == (1, 2 * GD, 2 * GH, 2 * GW) and y.dtype == torch.float32
    assert torch.equal(y, ref)


def test_gcnet_full_size_vs_oracle():
    """One 256x512 crop (feature maps 32 x 128 x 256, D = 96) through GCNetHotPath vs the CPU oracle: the gates of the
    small-size tests (kernel error vs the bf16-emulating oracle, format error vs fp32) at BASELINE config 3's extents;
    the numbers go to profiles/ through tools/parity_fullsize.py."""
    from dsmnet_b200.gcnet import GCNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    maxdisp = 192
    fL, fR, disp2 = O.synthetic_stereo_features(GH, GW, d_lo=5.0, d_hi=40.0, seed=6)
    cost = O.concat_volume(fL, fR, maxdisp // 2, "gc")
    params = O.gcnet_random_params(seed=23, calibrate_on=cost)
    ref = O.gcnet_hotpath(params, fL, fR, maxdisp)
    emu = O.gcnet_hotpath(params, fL, fR, maxdisp, operand_dtype=(torch.bfloat16, torch.bfloat16))
    del cost
    m = GCNetHotPath(maxdisp)
    m.load_state_dict({"layer3d." + k: v for k, v in params.items()}, strict=False)
    m = m.cuda().eval()
    with torch.no_grad():
        disp = m(fL.cuda(), fR.cuda()).cpu()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    assert disp.shape == ref.shape == (1, 1, 2 * GH, 2 * GW) and bool(torch.isfinite(disp).all())
    d_emu = float((disp - emu).abs().mean()); d_fmt = float((emu - ref).abs().mean()); d_ref = float((disp - ref).abs().mean())
    print("gcnet 256x512: mean |ours-emu| %.4f, |emu-ref| %.4f, |ours-ref| %.4f px" % (d_emu, d_fmt, d_ref))
    assert d_emu < max(0.5 * d_fmt, 0.02)
    assert d_ref < 2.0 * d_fmt + 0.02
