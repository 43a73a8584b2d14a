"""GPU (B200): the 2-D trunk kernels (dsm_conv2d_fwd on the tcgen05 implicit-GEMM kernel, dsm_conv2d_first_fwd, dsm_spp_fwd)
and the trunk plans against the oracle and the reference-made fixtures.

Tolerances (bf16 operands, fp32 accumulation, bf16 activation storage):
  * one layer vs the oracle with the same operand rounding: <= 2^-7 of the output scale (one bf16 rounding of the result);
  * whole trunk vs the oracle's same-format emulation: relative L2 <= 2.5 % (measured 1.6 % / 0.7 %: accumulation-order
    rounding flips through 56 / 19 layers);
  * whole trunk vs the reference's fp32 output: no farther than the emulation is (x1.25 + 0.5 %) — the bf16 format itself
    costs 2.7 % (PSMNet) / 1.4 % (GC-Net) on these random, BatchNorm-calibrated weights with a bf16 residual stream;
  * with trunk2d.DETERMINISTIC (one MMA issuer per accumulator in the row-sharing kernel) the trunk is bit-reproducible from
    run to run; the default (three issuers in rotation) differs in the accumulation order only."""
import pytest
import torch
import torch.nn.functional as F

import oracle.ops as O
from conftest import load_golden
from helpers import trunk_params_from_golden

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def l2rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / b.norm())


CASES = [
    # cin, cout, H, W, k, stride, dil, rim_in, rim_out, residual, relu
    (32, 32, 20, 37, 3, 1, 1, 1, 1, True, 1),
    (64, 64, 17, 50, 3, 1, 1, 2, 2, True, 0),
    (128, 128, 13, 41, 3, 1, 2, 2, 2, True, 0),
    (128, 128, 13, 41, 3, 1, 1, 2, 2, False, 1),
    (64, 128, 16, 33, 3, 1, 1, 2, 2, False, 1),
    (32, 64, 24, 46, 3, 2, 1, 1, 2, False, 1),
    (32, 64, 23, 45, 1, 2, 1, 1, 2, False, 0),
    (64, 128, 12, 30, 1, 1, 1, 2, 2, False, 0),
    (320, 128, 11, 29, 3, 1, 1, 2, 2, False, 1),
    (128, 32, 15, 31, 1, 1, 1, 2, 0, False, 0),
    # row-sharing kernel (conv2d_rs.cu): several bands and column tiles, ragged last band / tile, both dilations
    (64, 64, 37, 300, 3, 1, 1, 2, 2, True, 1),
    (64, 64, 21, 150, 3, 1, 2, 2, 2, True, 0),
    (32, 32, 50, 260, 3, 1, 1, 1, 1, True, 0),
    (32, 64, 19, 131, 3, 1, 1, 1, 2, False, 1),
    (64, 32, 11, 127, 3, 1, 2, 2, 1, False, 1),
    (64, 64, 1, 5, 3, 1, 1, 1, 1, False, 0),
    (128, 128, 23, 200, 3, 1, 2, 2, 2, True, 1),
    (128, 64, 9, 140, 3, 1, 1, 2, 2, True, 0),
    (64, 128, 30, 129, 3, 1, 1, 1, 1, False, 1),
]


@pytest.mark.parametrize("cin,cout,H,W,k,stride,dil,ri,ro,use_res,relu", CASES)
def test_conv2d_layer_vs_oracle(cin, cout, H, W, k, stride, dil, ri, ro, use_res, relu):
    from dsmnet_b200.trunk2d import FusedConv2d, PaddedImage
    from dsmnet_b200.conv3d import conv_timeouts
    torch.manual_seed(0)
    B = 2
    x = torch.randn(B, cin, H, W)
    conv = torch.nn.Conv2d(cin, cout, k, stride, padding=dil * (k // 2) if k > 1 else 0, dilation=dil, bias=False)
    conv.weight.data.normal_(0, (2.0 / (k * k * cin)) ** 0.5)
    bn = torch.nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
    sc, sh = O.fold_bn(cout, dict(weight=bn.weight.data, bias=bn.bias.data, running_mean=bn.running_mean, running_var=bn.running_var), None)
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if stride == 2 else (H, W)
    res = torch.randn(B, cout, Ho, Wo).to(BF).float() if use_res else None
    ref = O.conv2d_block(x, conv.weight.data, sc, sh, stride, dil, res, bool(relu), BF, None)
    layer = FusedConv2d(conv, bn, relu, "cuda")
    xin = PaddedImage.from_nchw(x.cuda(), ri)
    if ro == 0:                                                  # fp32 NCHW output mode
        out = torch.empty(B, cout, Ho, Wo, device="cuda")
        layer(xin, out)
        got = out.cpu()
        tol = 2e-5
    else:
        # write into a channel slice of a wider buffer and read the input from a slice too (ld > C paths)
        wide_in = PaddedImage.zeros(B, cin + 64, H, W, ri, "cuda")
        wide_in.view5()[..., 32:32 + cin] = xin.view5()
        wide_out = PaddedImage.zeros(B, cout + 32, Ho, Wo, ro, "cuda")
        rimg = PaddedImage.from_nchw(res.cuda(), ro) if use_res else None
        layer(wide_in, wide_out, x_c0=32, out_c0=16, residual=rimg)
        got = wide_out.to_nchw(16, cout).cpu()
        v = wide_out.view5()
        assert float(v[..., :16].abs().max()) == 0.0 and float(v[..., 16 + cout:].abs().max()) == 0.0      # neighbours untouched
        assert float(v[:, :ro].abs().max()) == 0.0 and float(v[:, :, :ro].abs().max()) == 0.0                # rim stays zero
        tol = 2.0 ** -7
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= tol * float(ref.abs().max())
    if ro != 0 and k == 3 and stride == 1 and cin <= 128 and cout <= 128:
        # these ran on the row-sharing kernel; the per-tile kernel (variant bit 2) must agree to one bf16 rounding
        layer.variant |= 4
        wide_out2 = PaddedImage.zeros(B, cout + 32, Ho, Wo, ro, "cuda")
        layer(wide_in, wide_out2, x_c0=32, out_c0=16, residual=rimg)
        got2 = wide_out2.to_nchw(16, cout).cpu()
        assert float((got - got2).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())


@pytest.mark.parametrize("k,H,W", [(3, 37, 61), (5, 40, 58)])
def test_first_conv_vs_torch(k, H, W):
    from dsmnet_b200.trunk2d import FirstConv2d, PaddedImage
    torch.manual_seed(1)
    img = torch.randn(2, 3, H, W)
    conv = torch.nn.Conv2d(3, 32, k, 2, k // 2, bias=(k == 5))
    bn = torch.nn.BatchNorm2d(32).eval()
    bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
    with torch.no_grad():
        ref = F.relu(bn(conv(img)))
    out = PaddedImage.zeros(2, 32, ref.shape[2], ref.shape[3], 1, "cuda")
    FirstConv2d(conv, bn, True, "cuda")(img.cuda(), out)
    got = out.to_nchw().cpu()
    assert float((got - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max()) + 1e-6      # fp32 arithmetic, one bf16 rounding


def test_spp_vs_oracle():
    from dsmnet_b200 import _lib
    from dsmnet_b200.trunk2d import PaddedImage
    torch.manual_seed(2)
    B, H, W = 2, 72, 136
    skip = torch.randn(B, 128, H, W).to(BF).float()
    w = torch.randn(4, 32, 128) * 0.1; sc = torch.rand(4, 32) + 0.5; sh = torch.randn(4, 32) * 0.3
    cat = PaddedImage.zeros(B, 320, H, W, 2, "cuda")
    cat.view5()[:, 2:2 + H, 2:2 + W, 64:192] = skip.permute(0, 2, 3, 1).to(BF).cuda()
    L = _lib.lib()
    ws = torch.empty(L.dsm_spp_workspace_bytes(B, H, W) // 4, device="cuda")
    wc, scc, shc = w.cuda(), sc.cuda(), sh.cuda()
    _lib.check(L.dsm_spp_fwd(cat.ptr(64), wc.data_ptr(), scc.data_ptr(), shc.data_ptr(), cat.ptr(0), B, H, W, 2, 320, 192, 1,
                             ws.data_ptr(), ws.numel() * 4, _lib.stream_ptr("cuda")), "dsm_spp_fwd")
    got = cat.to_nchw(192, 128).cpu()
    refs = []
    for i, k in enumerate((64, 32, 16, 8)):
        p = F.avg_pool2d(skip, (k, k), stride=(k, k))
        p = O.conv2d_block(p, w[i].view(32, 128, 1, 1), sc[i], sh[i], 1, 1, None, True, padding=1)
        refs.append(F.interpolate(p, (H, W), mode="bilinear", align_corners=True))
    ref = torch.cat((refs[3], refs[2], refs[1], refs[0]), 1)
    assert float((got - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())
    assert torch.equal(cat.to_nchw(64, 128).cpu(), skip)                                      # the skip slice is untouched


def test_psmnet_trunk_vs_reference_golden():
    from dsmnet_b200.psmnet import feature_extraction
    from dsmnet_b200.conv3d import conv_timeouts
    g = load_golden("psmnet_trunk")
    m = feature_extraction().eval()
    p = trunk_params_from_golden(g, m)
    m.load_state_dict(p, strict=False)
    with torch.no_grad():
        emu = O.psmnet_feature_extraction(p, g["x"], operand_dtype=(BF, BF))
        out = m.cuda()(g["x"].cuda()).cpu()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    e_emu, e_ref, e_fmt = l2rel(out, emu), l2rel(out, g["out"]), l2rel(emu, g["out"])
    print("psmnet trunk: rel L2 ours-emu %.4f, ours-ref %.4f, emu-ref %.4f" % (e_emu, e_ref, e_fmt))
    assert out.shape == g["out"].shape
    assert e_emu <= 0.025 and e_ref <= 1.25 * e_fmt + 0.005
    # the single-issuer mode of the row-sharing kernel makes the trunk bit-reproducible from run to run
    from dsmnet_b200 import trunk2d
    trunk2d.DETERMINISTIC = True
    try:
        m.__dict__.pop("_dsm_plan", None)
        with torch.no_grad():
            a = m(g["x"].cuda()).cpu()
            assert torch.equal(m(g["x"].cuda()).cpu(), a)
        assert l2rel(a, out) <= 0.02
    finally:
        trunk2d.DETERMINISTIC = False
        m.__dict__.pop("_dsm_plan", None)


def test_gcnet_trunk_vs_reference_golden():
    from dsmnet_b200.gcnet import feature2d
    g = load_golden("gcnet_trunk")
    m = feature2d(32).eval()
    p = trunk_params_from_golden(g, m)
    m.load_state_dict(p, strict=False)
    with torch.no_grad():
        emu = O.gcnet_feature2d(p, g["x"], operand_dtype=(BF, BF))
        out = m.cuda()(g["x"].cuda()).cpu()
    e_emu, e_ref, e_fmt = l2rel(out, emu), l2rel(out, g["out"]), l2rel(emu, g["out"])
    print("gcnet trunk: rel L2 ours-emu %.4f, ours-ref %.4f, emu-ref %.4f" % (e_emu, e_ref, e_fmt))
    assert out.shape == g["out"].shape
    assert e_emu <= 0.025 and e_ref <= 1.25 * e_fmt + 0.005


def test_psmnet_whole_model_cuda_matches_its_own_stock_graph():
    """PSMNet(left, right) on CUDA under no_grad (trunk plan + hot path) vs the same module's stock-PyTorch trunk feeding the
    same hot path: the disparity maps differ only by the trunk's bf16 format."""
    from dsmnet_b200.psmnet import PSMNet, PSMNetHotPath
    torch.manual_seed(3)
    g = load_golden("psmnet_trunk")
    m = PSMNet(64).eval()
    p = trunk_params_from_golden(g, m.feature_extraction)
    m.feature_extraction.load_state_dict(p, strict=False)
    m.load_state_dict(O.psmnet_matcher_params(seed=21), strict=False)      # a 3-D stack that matches: not a chaotic amplifier
    m = m.cuda()
    left = g["x"].cuda(); right = torch.roll(left, -6, dims=3)
    with torch.no_grad():
        _, preds = m(left, right)
        fe = m.feature_extraction
        with torch.enable_grad():                                  # forces the stock graph
            fl = fe(left).detach(); fr = fe(right).detach()
        ref = [r.clone() for r in PSMNetHotPath.forward(m, fl, fr, (left.size(2), left.size(3)))]
        _, preds2 = m(left, right)
    for a, b, c in zip(preds, ref, preds2):
        assert a.shape == b.shape == (1, 256, 320)
        # two runs of the SAME path differ by accumulation-order rounding flips of the 3-D plane-sharing kernel (three MMA
        # issuers per accumulator); the trunk's bf16 format must not add more than a few times that on this random network
        noise = float((a - c).abs().mean()); d = float((a - b).abs().mean())
        print("whole model: mean |d(trunk plan) - d(stock trunk)| %.4f px; run-to-run %.4f px" % (d, noise))
        assert d < 2.0 and noise < 0.3
