"""GPU (B200): the tcgen05 conv3d blocks and the PSMNet hot path against the oracle / golden vectors.

bf16 tolerance: conv operands are rounded to bf16, accumulation is fp32, activations are stored
in bf16.  Layer level: against the oracle run with bf16-rounded operands, error <= 2^-7 of the
output scale (one bf16 rounding of the result).  Path level, two gates:
  (1) kernel correctness: against the oracle evaluated with the SAME rounding points (bf16
      operands, bf16 activation storage, fp32 accumulation / classifier / soft-argmin).  What is
      left is rounding-flip noise (fp32 accumulation order decides which way a stored activation
      rounds, and this random-weight network amplifies every flip), so the gate is relative: the
      CUDA path must sit closer to its own CPU emulation than that emulation sits to fp32, the
      low-resolution costs must agree to 2% in L2, and mean |disparity delta| < 0.1 px;
  (2) format error: against the fp32 oracle / the reference's own fp32 modules the delta is the
      intrinsic cost of bf16 operands on this random-weight network (logit std ~8): measured
      0.06-0.16 px mean |delta|, the same as the CPU bf16 emulation shows; bounded at 0.25 px and
      recorded in DESIGN.md.  The north_star's metric — the change of the mean end-point error
      against a ground-truth disparity map ("mean EPE delta < 0.01 px") — is asserted at the
      BASELINE size in test_gpu_fullsize.py, where there are enough pixels for a mean."""
import pytest
import torch

BF16 = (torch.bfloat16, torch.bfloat16)


def l2rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / b.norm())
import oracle.ops as O
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def run_layer(x, weight, scale, shift, stride, transposed, residual, relu, variant=0):
    from dsmnet_b200.conv3d import FusedConv3d, conv_timeouts
    from dsmnet_b200.volume_layout import PaddedVolume

    class BN:  # minimal stand-in carrying folded scale/shift
        pass
    layer = FusedConv3d(weight.cuda(), None, None, stride, transposed, relu, variant=variant)
    if scale is not None:
        layer.set_affine(scale, shift)
    xv = PaddedVolume.from_ncdhw(x.cuda())
    rv = None
    if residual is not None:
        rv = residual.cuda().contiguous() if layer.cout == 1 else PaddedVolume.from_ncdhw(residual.cuda())
        if layer.cout == 1:
            rv = rv.squeeze(1)
    y = layer(xv, None, rv)
    assert conv_timeouts() == 0
    return y.unsqueeze(1).cpu() if layer.cout == 1 else y.to_ncdhw().cpu()


CASES = [
    # cin, cout, (D,H,W), stride, transposed, residual, relu
    (32, 32, (4, 6, 20), 1, False, True, True),
    (64, 32, (5, 7, 19), 1, False, False, True),
    (64, 64, (6, 9, 21), 1, False, True, True),
    (32, 1, (4, 6, 20), 1, False, True, False),
    (32, 64, (8, 12, 40), 2, False, False, True),
    (64, 64, (7, 11, 37), 2, False, False, True),
    (64, 64, (3, 5, 10), 2, True, True, True),
    (64, 32, (3, 5, 10), 2, True, True, False),
    (32, 1, (3, 5, 10), 2, True, False, False),
    (128, 128, (3, 4, 9), 1, False, False, True),
    (64, 128, (6, 8, 18), 2, False, False, True),
    (128, 64, (3, 4, 9), 2, True, True, True),
]


@pytest.mark.parametrize("cin,cout,dims,stride,transposed,use_res,relu", CASES)
def test_conv_layer_vs_oracle(cin, cout, dims, stride, transposed, use_res, relu):
    torch.manual_seed(0)
    B = 2
    x = torch.randn(B, cin, *dims)
    w = torch.randn(cin, cout, 3, 3, 3) if transposed else torch.randn(cout, cin, 3, 3, 3)
    w = w * (2.0 / (27 * cin)) ** 0.5
    scale = torch.rand(cout) + 0.5; shift = torch.randn(cout) * 0.3
    ref0 = O.conv3d_block(x, w, scale, shift, stride, transposed, None, False, torch.bfloat16)
    residual = torch.randn_like(ref0) if use_res else None
    if residual is not None and cout != 1:
        residual = residual.to(torch.bfloat16).float()
    ref = O.conv3d_block(x, w, scale, shift, stride, transposed, residual, relu, torch.bfloat16)
    out = run_layer(x, w, scale, shift, stride, transposed, residual, relu)
    assert out.shape == ref.shape
    tol = (2e-5 if cout == 1 else 2.0 ** -7) * float(ref.abs().max())
    assert float((out - ref).abs().max()) <= tol


@pytest.mark.parametrize("name,transposed", [("gc_conv_s2", False), ("gc_deconv", True)])
def test_gc_layers_golden(name, transposed):
    g = load_golden(name)
    bn = {"weight": g["bn_weight"], "bias": g["bn_bias"], "running_mean": g["bn_mean"], "running_var": g["bn_var"]}
    scale, shift = O.fold_bn(32, bn, g["bias"])
    out = run_layer(g["x"], g["weight"], scale, shift, 2, transposed, None, True)
    assert out.shape == g["y"].shape
    assert float((out - g["y"]).abs().max()) <= 2.0 ** -6 * float(g["y"].abs().max())   # bf16 operands + bf16 output


def test_crop_semantics_odd_sizes():
    """myadd_3d crop-to-min (stackhourglass.py:10-20): deconv of an odd-sized skip connection."""
    torch.manual_seed(1)
    x = torch.randn(1, 64, 3, 4, 6)                       # deconv -> 6 x 8 x 12
    w = torch.randn(64, 32, 3, 3, 3) * 0.05
    res = torch.randn(1, 32, 5, 7, 11).to(torch.bfloat16).float()   # smaller skip -> result is 5 x 7 x 11
    ref = O.conv3d_block(x, w, None, None, 2, True, res, True, torch.bfloat16)
    out = run_layer(x, w, None, None, 2, True, res, True)
    assert out.shape == ref.shape == (1, 32, 5, 7, 11)
    assert float((out - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())


def _hotpath_module(params, maxdisp):
    from dsmnet_b200.psmnet import PSMNetHotPath
    m = PSMNetHotPath(maxdisp)
    missing = m.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing.missing_keys
    return m.cuda().eval()


def test_psmnet_hotpath_golden():
    """the reference's own PSMNet 3-D stack + heads (fixture from its modules) vs the CUDA path."""
    import hashlib
    g = load_golden("psmnet_hotpath")
    cost = O.concat_volume(g["fL"], g["fR"], g["maxdisp"] // 4, "psm")
    params = O.psmnet_random_params(seed=g["seed"], calibrate_on=cost)
    h = hashlib.sha256()
    for k in sorted(params):
        if params[k].dim() == 5:
            h.update(k.encode()); h.update(params[k].numpy().tobytes())
    assert h.hexdigest() == g["params_sha256"], "synthetic weights are not the ones the fixture was made with"
    m = _hotpath_module(params, g["maxdisp"])
    with torch.no_grad():
        c1, c2, c3 = m.aggregate(g["fL"].cuda(), g["fR"].cuda())
        preds = m(g["fL"].cuda(), g["fR"].cuda(), (g["H"], g["W"]))
    ecost = O.psmnet_aggregate(params, cost.to(torch.bfloat16).float(), BF16)
    for mine, ref, e in zip((c1, c2, c3), (g["cost1"], g["cost2"], g["cost3"]), ecost):
        assert l2rel(mine.unsqueeze(1), ref) < 5e-2                      # gate (2) on the low-res costs
        assert l2rel(mine.unsqueeze(1), e) < 2e-2                        # gate (1)
    emu = O.psmnet_hotpath(params, g["fL"], g["fR"], g["maxdisp"], (g["H"], g["W"]), operand_dtype=BF16)
    for mine, ref, e in zip(preds, (g["pred3"], g["pred2"], g["pred1"]), emu):
        mine = mine.cpu()
        d_emu = float((mine - e).abs().mean()); d_ref = float((mine - ref).abs().mean())
        assert d_emu < 0.1 and d_emu < float((e - ref).abs().mean())     # gate (1)
        assert d_ref < 0.25                                              # gate (2): bf16 format error vs the reference


@pytest.mark.parametrize("B,H,W,maxdisp", [(1, 48, 96, 48), (2, 24, 40, 32)])
def test_psmnet_hotpath_vs_oracle(B, H, W, maxdisp):
    torch.manual_seed(5)
    fL = torch.randn(B, 32, H // 4, W // 4); fR = torch.randn(B, 32, H // 4, W // 4)
    cost = O.concat_volume(fL, fR, maxdisp // 4, "psm")
    params = O.psmnet_random_params(seed=11, calibrate_on=cost)
    ref = O.psmnet_hotpath(params, fL, fR, maxdisp, (H, W))
    emu = O.psmnet_hotpath(params, fL, fR, maxdisp, (H, W), operand_dtype=BF16)
    m = _hotpath_module(params, maxdisp)
    with torch.no_grad():
        preds = m(fL.cuda(), fR.cuda(), (H, W))
    from dsmnet_b200.conv3d import conv_timeouts
    assert conv_timeouts() == 0
    for mine, r, e in zip(preds, ref, emu):
        assert mine.shape == r.shape == (B, H, W)
        d_emu = float((mine.cpu() - e).abs().mean())
        assert d_emu < 0.1 and d_emu < float((e - r).abs().mean())       # gate (1)
        assert float((mine.cpu() - r).abs().mean()) < 0.25               # gate (2)


def _gc_module(params, maxdisp):
    from dsmnet_b200.gcnet import GCNetHotPath
    m = GCNetHotPath(maxdisp)
    missing = m.load_state_dict({"layer3d." + k: v for k, v in params.items()}, strict=False)
    assert not missing.unexpected_keys
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing.missing_keys
    return m.cuda().eval()


def test_gcnet_hotpath_golden():
    """GC-Net's 3-D path (gcnet.py:65-111 run by the reference itself) vs the CUDA path."""
    g = load_golden("gcnet_hotpath")
    cost = O.concat_volume(g["fL"], g["fR"], g["maxdisp"] // 2, "gc")
    params = O.gcnet_random_params(seed=g["seed"], calibrate_on=cost)
    m = _gc_module(params, g["maxdisp"])
    with torch.no_grad():
        disp = m(g["fL"].cuda(), g["fR"].cuda()).cpu()
    emu = O.gcnet_hotpath(params, g["fL"], g["fR"], g["maxdisp"], operand_dtype=BF16)
    assert disp.shape == g["disp"].shape
    d_emu = float((disp - emu).abs().mean()); d_fmt = float((emu - g["disp"]).abs().mean()); d_ref = float((disp - g["disp"]).abs().mean())
    print("gcnet: mean |ours-emu| %.4f, |emu-ref| %.4f, |ours-ref| %.4f px" % (d_emu, d_fmt, d_ref))
    assert d_emu < max(0.5 * d_fmt, 0.02)             # gate (1): the kernels add little beyond the operand format
    assert d_ref < 2.0 * d_fmt + 0.02                 # gate (2)


@pytest.mark.parametrize("B,H,W,maxdisp", [(1, 20, 28, 24), (2, 16, 16, 16)])
def test_gcnet_hotpath_vs_oracle(B, H, W, maxdisp):
    """odd level sizes exercise the myAdd3d crop of the skip connections (util_fun.py:41-51)"""
    torch.manual_seed(6)
    fL = torch.randn(B, 32, H, W); fR = torch.randn(B, 32, H, W)
    cost = O.concat_volume(fL, fR, maxdisp // 2, "gc")
    params = O.gcnet_random_params(seed=13, calibrate_on=cost)
    ref = O.gcnet_hotpath(params, fL, fR, maxdisp)
    emu = O.gcnet_hotpath(params, fL, fR, maxdisp, operand_dtype=BF16)
    m = _gc_module(params, maxdisp)
    with torch.no_grad():
        disp = m(fL.cuda(), fR.cuda()).cpu()
    from dsmnet_b200.conv3d import conv_timeouts
    assert conv_timeouts() == 0
    assert disp.shape == ref.shape
    d_emu = float((disp - emu).abs().mean()); d_fmt = float((emu - ref).abs().mean())
    print("gcnet %s: mean |ours-emu| %.4f, |emu-ref| %.4f px" % ((B, H, W, maxdisp), d_emu, d_fmt))
    assert d_emu < max(0.5 * d_fmt, 0.02)
    assert float((disp - ref).abs().mean()) < 2.0 * d_fmt + 0.02


def test_gcnet_training_step_gradients():
    """train-mode GC-Net 3-D stack (BASELINE config 3 is fwd+bwd): loss and gradients vs torch autograd of the oracle"""
    from dsmnet_b200.gcnet import GCNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    torch.manual_seed(8)
    B, h, w, maxdisp = 2, 16, 32, 32                       # D = 16 -> 8 -> 4 -> 2 -> 1
    fL = torch.randn(B, 32, h, w); fR = torch.randn(B, 32, h, w)
    params = O.gcnet_random_params(seed=19)
    gt = torch.rand(B, 1, 2 * h, 2 * w) * maxdisp * 0.5
    req = lambda k, v: v.dim() == 5 or k.endswith(".bias") or k.endswith(".1.weight")
    pr = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    a = fL.clone().requires_grad_(); b = fR.clone().requires_grad_()
    lref = (O.gcnet_hotpath_train(pr, a, b, maxdisp) - gt).abs().mean(); lref.backward()
    pe = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    ae = fL.clone().requires_grad_(); be = fR.clone().requires_grad_()
    lemu = (O.gcnet_hotpath_train(pe, ae, be, maxdisp, operand_dtype=torch.bfloat16) - gt).abs().mean(); lemu.backward()
    m = GCNetHotPath(maxdisp)
    m.load_state_dict({"layer3d." + k: v for k, v in params.items()}, strict=False)
    m = m.cuda().train()
    x = fL.cuda().requires_grad_(); y = fR.cuda().requires_grad_()
    loss = (m(x, y) - gt.cuda()).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    print("gcnet loss: ours %.5f, oracle fp32 %.5f, bf16-emulation %.5f" % (loss.item(), lref.item(), lemu.item()))
    assert abs(loss.item() - lref.item()) < 0.03 * abs(lref.item())
    cos = lambda u, v: float(torch.nn.functional.cosine_similarity(u.flatten().double(), v.flatten().double(), dim=0))
    named = dict(m.layer3d.named_parameters())
    checks = [("fL", x.grad.cpu(), a.grad, ae.grad), ("fR", y.grad.cpu(), b.grad, be.grad)]
    for k in ("l19.0.weight", "l21.0.weight", "l22.0.weight", "l30.0.weight", "l32.0.weight", "l33.0.weight", "l36.0.weight",
              "l37.weight", "l35.1.weight", "l34.1.bias"):     # (l37.bias shifts every logit alike and conv biases feed a batch-stat BN: analytically zero gradients)
        checks.append((k, named[k].grad.cpu(), pr[k].grad, pe[k].grad))
    for name, mine, ref, emu in checks:
        c_ref, c_fmt = cos(mine, ref), cos(emu, ref)
        r_ref, r_emu = float(mine.norm() / ref.norm()), float(mine.norm() / emu.norm())
        print("grad %-16s cos(ours,fp32) %.4f  cos(emu,fp32) %.4f  |ours|/|ref| %.3f  |ours|/|emu| %.3f" % (name, c_ref, c_fmt, r_ref, r_emu))
        # the bottom of the encoder normalises over a handful of voxels: the bf16 format alone moves those gradients by
        # ~15 % (emu vs fp32), so the magnitude gate is tight against the emulation and loose against fp32
        assert c_ref > 0.9 and c_ref > c_fmt - 0.03 and 0.7 < r_ref < 1.3 and 0.9 < r_emu < 1.1


def test_psmnet_training_step_gradients():
    """train-mode PSMNet 3-D stack (batch-stat BatchNorm): loss and gradients w.r.t. the feature maps and the
    parameters vs torch autograd of the oracle's fp32 restatement (stackhourglass.py:123-168 under model.train())."""
    from dsmnet_b200.psmnet import PSMNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    torch.manual_seed(3)
    B, h, w, maxdisp = 2, 8, 16, 32
    fL = torch.randn(B, 32, h, w); fR = torch.randn(B, 32, h, w)
    params = O.psmnet_random_params(seed=17)
    gt = torch.rand(B, 4 * h, 4 * w) * maxdisp * 0.5

    def loss_of(preds):
        return sum(wt * (p - gt).abs().mean() for wt, p in zip((1.0, 0.7, 0.5), preds))     # the reference's weighted L1 pyramid

    # oracle, fp32
    pr = {k: v.clone().requires_grad_(v.dim() == 5 or k.endswith(".1.weight") or k.endswith(".1.bias")) for k, v in params.items()}
    a = fL.clone().requires_grad_(); b = fR.clone().requires_grad_()
    lref = loss_of(O.psmnet_hotpath_train(pr, a, b, maxdisp, (4 * h, 4 * w)))
    lref.backward()
    # oracle with bf16 operand / storage emulation (straight-through): the format-error yardstick
    pe = {k: v.clone().requires_grad_(v.dim() == 5 or k.endswith(".1.weight") or k.endswith(".1.bias")) for k, v in params.items()}
    ae = fL.clone().requires_grad_(); be = fR.clone().requires_grad_()
    lemu = loss_of(O.psmnet_hotpath_train(pe, ae, be, maxdisp, (4 * h, 4 * w), operand_dtype=torch.bfloat16))
    lemu.backward()
    # CUDA path
    m = PSMNetHotPath(maxdisp)
    m.load_state_dict(params, strict=False)
    m = m.cuda().train()
    x = fL.cuda().requires_grad_(); y = fR.cuda().requires_grad_()
    gtc = gt.cuda()
    preds = m(x, y, (4 * h, 4 * w))
    loss = sum(wt * (p - gtc).abs().mean() for wt, p in zip((1.0, 0.7, 0.5), preds))
    loss.backward()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    print("loss: ours %.5f, oracle fp32 %.5f, oracle bf16-emulation %.5f" % (float(loss), float(lref), float(lemu)))
    assert abs(float(loss) - float(lref)) < 0.02 * abs(float(lref))

    def cos(u, v):
        return float(torch.nn.functional.cosine_similarity(u.flatten().double(), v.flatten().double(), dim=0))

    named = dict(m.named_parameters())
    checks = [("fL", x.grad.cpu(), a.grad, ae.grad), ("fR", y.grad.cpu(), b.grad, be.grad)]
    for k in ("dres0.0.0.weight", "dres0.2.0.weight", "dres1.2.0.weight", "dres2.conv1.0.0.weight", "dres2.conv2.0.weight",
              "dres3.conv5.0.weight", "dres4.conv6.0.weight", "classif1.0.0.weight", "classif3.2.weight",
              "dres2.conv2.1.weight", "dres4.conv6.1.bias"):
        checks.append((k, named[k].grad.cpu(), pr[k].grad, pe[k].grad))
    for name, mine, ref, emu in checks:
        c_ref, c_emu, c_fmt = cos(mine, ref), cos(mine, emu), cos(emu, ref)
        print("grad %-26s cos(ours,fp32) %.4f  cos(ours,emu) %.4f  cos(emu,fp32) %.4f  |ours|/|ref| %.3f" %
              (name, c_ref, c_emu, c_fmt, float(mine.norm() / ref.norm())))
        # direction and magnitude of every gradient; the yardstick is what the bf16 operand/storage format alone does to
        # the fp32 gradient of this random, batch-normalised (hence noise-amplifying) network: cos(emu, fp32) ~ 0.97
        assert c_ref > 0.9 and 0.85 < float(mine.norm() / ref.norm()) < 1.15
        assert c_ref > c_fmt - 0.03                                             # no worse than the bf16 format itself


@pytest.mark.parametrize("cout,cin", [(32, 64), (1, 32), (64, 32), (128, 64)])
def test_pack_weight_kernel_matches_host_packing(cout, cin):
    """dsm_pack_weight (one launch) is bit-identical to the host-side permute/cat/cast packing in all three modes"""
    from dsmnet_b200.conv3d import pack_weight, pack_weight_device
    torch.manual_seed(9)
    w = torch.randn(cout, cin, 3, 3, 3)
    assert torch.equal(pack_weight_device(w.cuda(), 0).cpu(), pack_weight(w, False))                 # Conv3d
    wt = torch.randn(cin, cout, 3, 3, 3)
    assert torch.equal(pack_weight_device(wt.cuda(), 1).cpu(), pack_weight(wt, True))                # ConvTranspose3d
    if cout >= 16:
        ref = pack_weight(w.flip(2, 3, 4).transpose(0, 1).contiguous(), False)                        # stride-1 dgrad filter
        assert torch.equal(pack_weight_device(w.cuda(), 2).cpu(), ref)


@pytest.mark.parametrize("mode", ["psm", "gc", "gc_right"])
@pytest.mark.parametrize("B,D,H,W", [(2, 5, 7, 19), (1, 12, 9, 150), (1, 6, 20, 10)])
def test_fused_volume_conv_equals_volume_then_conv(mode, B, D, H, W):
    """dsm_conv3d_volume_fwd (the volume is never written) vs dsm_concat_volume_fwd + dsm_conv3d_fwd on the same weights, and
    vs the oracle: every mode, batch > 1, D larger than the rows of a tile, W < 16 (several image rows inside one tile)"""
    from dsmnet_b200.conv3d import FusedConv3d, conv_from_features, pack_features_nhwc, conv_timeouts
    from dsmnet_b200.cost_volume import concat_volume
    torch.manual_seed(21)
    fL = torch.randn(B, 32, H, W); fR = torch.randn(B, 32, H, W)
    w = torch.randn(32, 64, 3, 3, 3) * (2.0 / (27 * 64)) ** 0.5
    scale = torch.rand(32) + 0.5; shift = torch.randn(32) * 0.3
    layer = FusedConv3d(w.cuda(), None, None, 1, False, 1)
    layer.set_affine(scale, shift)
    fused = conv_from_features(layer, pack_features_nhwc(fL.cuda()), pack_features_nhwc(fR.cuda()), D, mode).to_ncdhw().cpu()
    plain = layer(concat_volume(fL.cuda(), fR.cuda(), D, mode, padded_bf16=True)).to_ncdhw().cpu()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    vol = O.concat_volume(fL, fR, D, mode)
    ref = O.conv3d_block(vol, w, scale, shift, 1, False, None, True, torch.bfloat16)
    tol = 2.0 ** -7 * float(ref.abs().max())
    assert float((fused - ref).abs().max()) <= tol
    # same operands, same MMAs: only the order in which the three issuing warps reach the accumulator may differ
    assert float((fused - plain).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())
    assert float((fused != plain).float().mean()) < 0.01


def test_fused_volume_conv_full_size():
    """BASELINE size (volume 64x48x96x312): the fused first layer against the materialised route"""
    from dsmnet_b200.conv3d import FusedConv3d, conv_from_features, pack_features_nhwc, conv_timeouts
    from dsmnet_b200.cost_volume import concat_volume
    torch.manual_seed(22)
    fL = torch.randn(1, 32, 96, 312, device="cuda"); fR = torch.randn(1, 32, 96, 312, device="cuda")
    w = torch.randn(32, 64, 3, 3, 3) * (2.0 / (27 * 64)) ** 0.5
    layer = FusedConv3d(w.cuda(), None, None, 1, False, 1)
    fused = conv_from_features(layer, pack_features_nhwc(fL), pack_features_nhwc(fR), 48, "psm").to_ncdhw()
    plain = layer(concat_volume(fL, fR, 48, "psm", padded_bf16=True)).to_ncdhw()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    assert float((fused - plain).abs().max()) <= 2.0 ** -7 * float(plain.abs().max())       # one bf16 ulp at the top binade
    assert float((fused != plain).float().mean()) < 0.01


@pytest.mark.parametrize("kwfold", [True, False])
@pytest.mark.parametrize("dims,B", [((4, 6, 20), 2), ((9, 13, 131), 1), ((3, 40, 7), 1)])
def test_single_channel_conv_kwfold_and_plain(kwfold, dims, B, monkeypatch):
    """classif*.2 (Conv3d 32 -> 1, fp32 output, cumulative residual): the kw-fold mode (three kw taps as output columns, tiles
    of 120 positions, shuffled column sum) and the plain 32 -> 16 mapping against the oracle — rows longer and shorter than a
    30-position quadrant, several image rows per tile"""
    import dsmnet_b200.conv3d as C3
    monkeypatch.setattr(C3, "KWFOLD", kwfold)
    torch.manual_seed(31)
    x = torch.randn(B, 32, *dims)
    w = torch.randn(1, 32, 3, 3, 3) * (2.0 / (27 * 32)) ** 0.5
    res = torch.randn(B, 1, *dims)
    ref = O.conv3d_block(x, w, None, None, 1, False, res, False, torch.bfloat16)
    out = run_layer(x, w, None, None, 1, False, res, False)
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
