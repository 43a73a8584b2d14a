"""GPU (B200): the TRAINING path of the 3-D stacks against (1) the reference's own modules under .train() (frozen by
tests/golden/make_golden_train.py: loss, predictions, gradients, running statistics) and (2) the oracle evaluated with the
kernels' number format (bf16 operands, bf16 activation and activation-gradient storage, fp32 accumulation / statistics /
weight gradients).

Gates, per gradient (input features and ten parameters per model), tied to the emulation as VERDICT r01 asked:
  * norm:      | |ours| / |emulation| - 1 | <= 5 %
  * direction: 1 - cos(ours, emulation) <= max(0.005, 1.25 * (1 - cos(emulation, fp32 reference)))
    i.e. the kernels sit about as close to their own number-format emulation as that emulation sits to fp32.  A fixed
    cos >= 0.995 does not hold for the deepest gradients of these fixtures and cannot: with batch-statistics BatchNorm over
    a few hundred voxels and random weights, the fp32 accumulation ORDER decides which way an activation rounds to bf16,
    a ReLU mask flips, and the network amplifies it.  The CUDA path itself is not bit-reproducible from run to run (three
    MMA issuers accumulate into one TMEM block in whatever order they get there), and two of ITS runs differ by as much:
    measured on B200 (profiles/r02_parity_train.txt) cos(ours, emulation) for the deepest gradient (fL) was 0.988 in one
    run and 0.973 in the next, with cos(emulation, fp32) = 0.968.  The shallow gradients (classifiers, l36 / l37, the
    BatchNorm affine of the last layers) meet 0.995 in every run; the layer-level backward tests
    (tests/test_gpu_conv3d_bwd.py, tests/test_gpu_bnact.py) are the tight, chaos-free kernel checks.
  * against the reference's fp32 gradients: cos(ours, ref) >= cos(emulation, ref) - 0.02.
The odd-sized cases run the cropped skip adds (BatchNorm statistics over the uncropped deconv output, crop at the add)
entirely on the fused kernels; the running statistics after the step are compared with the reference's."""
import pytest
import torch

import oracle.ops as O
from conftest import load_golden
from helpers import GC_TRAIN_GRADS, PSM_TRAIN_GRADS, cosine, golden_grad, psm_train_loss

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _report(tag, name, mine, ref, emu):
    c_ref, c_emu, c_fmt = cosine(mine, ref), cosine(mine, emu), cosine(emu, ref)
    r_ref, r_emu = float(mine.norm() / ref.norm()), float(mine.norm() / emu.norm())
    print("%s grad %-24s cos(ours,emu) %.5f  cos(ours,ref) %.5f  cos(emu,ref) %.5f  |ours|/|emu| %.4f  |ours|/|ref| %.4f" %
          (tag, name, c_emu, c_ref, c_fmt, r_emu, r_ref))
    return c_ref, c_emu, c_fmt, r_ref, r_emu


def _gates(vals, failures, what):
    c_ref, c_emu, c_fmt, r_ref, r_emu = vals
    if abs(r_emu - 1.0) > 0.05:
        failures.append("%s: |ours|/|emu| = %.4f" % (what, r_emu))
    if (1.0 - c_emu) > max(0.005, 1.25 * (1.0 - c_fmt)):
        failures.append("%s: cos(ours,emu) = %.5f with cos(emu,ref) = %.5f" % (what, c_emu, c_fmt))
    if c_ref < c_fmt - 0.02:
        failures.append("%s: cos(ours,ref) = %.5f < cos(emu,ref) = %.5f - 0.02" % (what, c_ref, c_fmt))


@pytest.mark.parametrize("name", ["psmnet_train", "psmnet_train_odd"])
def test_psmnet_training_vs_reference_golden(name):
    from dsmnet_b200.psmnet import PSMNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    g = load_golden(name)
    params = O.psmnet_random_params(seed=g["seed"])
    H, W = g["gt"].shape[-2:]
    req = lambda k, v: v.dim() == 5 or k.endswith(".1.weight") or k.endswith(".1.bias")
    pe = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    ae = g["fL"].clone().requires_grad_(); be = g["fR"].clone().requires_grad_()
    pemu = O.psmnet_hotpath_train(pe, ae, be, g["maxdisp"], (H, W), operand_dtype=BF, grad_dtype=BF)
    lemu = psm_train_loss(pemu, g["gt"]); lemu.backward()

    m = PSMNetHotPath(g["maxdisp"])
    m.load_state_dict(params, strict=False)
    m = m.cuda().train()
    x = g["fL"].cuda().requires_grad_(); y = g["fR"].cuda().requires_grad_()
    preds = m(x, y, (H, W))
    loss = psm_train_loss(preds, g["gt"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    print("%s loss: ours %.5f  reference %.5f  emulation %.5f" % (name, float(loss), float(g["loss"]), float(lemu)))
    assert abs(float(loss) - float(lemu)) < 0.01 * abs(float(lemu))
    assert abs(float(loss) - float(g["loss"])) < 0.03 * abs(float(g["loss"]))
    for mine, ref, e in zip(preds, (g["pred3"], g["pred2"], g["pred1"]), pemu):
        d_emu = float((mine.detach().cpu() - e.detach()).abs().mean()); d_ref = float((mine.detach().cpu() - ref).abs().mean())
        print("%s pred: mean |ours-emu| %.4f px, |ours-ref| %.4f px" % (name, d_emu, d_ref))
        assert d_emu < 0.1 and d_ref < 0.25                  # rounding-flip noise / bf16 format error, as in the inference tests
    named = dict(m.named_parameters())
    failures = []
    _gates(_report(name, "fL", x.grad.cpu(), g["gL"], ae.grad), failures, "fL")
    _gates(_report(name, "fR", y.grad.cpu(), g["gR"], be.grad), failures, "fR")
    for k in PSM_TRAIN_GRADS:
        mine, ref = golden_grad(g, k, named[k].grad)
        emu = golden_grad(g, k, pe[k].grad)[0]
        _gates(_report(name, k, mine, ref, emu), failures, k)
    assert not failures, failures
    # running statistics after one step (momentum 0.1 from (0, 1)): BatchNorm saw the UNCROPPED conv5 output
    sd = m.state_dict()
    for mine, ref in ((sd["dres0.0.1.running_mean"], g["rm_dres0_0"]), (sd["dres0.0.1.running_var"], g["rv_dres0_0"]),
                      (sd["dres3.conv5.1.running_mean"], g["rm_conv5"]), (sd["dres3.conv5.1.running_var"], g["rv_conv5"])):
        assert float((mine.cpu() - ref).abs().max()) < 2e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("name", ["gcnet_train", "gcnet_train_odd"])
def test_gcnet_training_vs_reference_golden(name):
    from dsmnet_b200.gcnet import GCNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    g = load_golden(name)
    params = O.gcnet_random_params(seed=g["seed"])
    req = lambda k, v: v.dim() == 5 or k.endswith(".bias") or k.endswith(".1.weight")
    pe = {k: v.clone().requires_grad_(req(k, v)) for k, v in params.items()}
    ae = g["fL"].clone().requires_grad_(); be = g["fR"].clone().requires_grad_()
    demu = O.gcnet_hotpath_train(pe, ae, be, g["maxdisp"], operand_dtype=BF, grad_dtype=BF)
    gt = g["gt"][:, :, :demu.shape[2], :demu.shape[3]]
    lemu = (demu - gt).abs().mean(); lemu.backward()

    m = GCNetHotPath(g["maxdisp"])
    m.load_state_dict({"layer3d." + k: v for k, v in params.items()}, strict=False)
    m = m.cuda().train()
    x = g["fL"].cuda().requires_grad_(); y = g["fR"].cuda().requires_grad_()
    disp = m(x, y)
    assert disp.shape == g["disp"].shape
    loss = (disp - gt.cuda()).abs().mean()
    loss.backward()
    torch.cuda.synchronize()
    assert conv_timeouts() == 0
    print("%s loss: ours %.5f  reference %.5f  emulation %.5f" % (name, float(loss), float(g["loss"]), float(lemu)))
    assert abs(float(loss) - float(lemu)) < 0.01 * abs(float(lemu))
    assert abs(float(loss) - float(g["loss"])) < 0.03 * abs(float(g["loss"]))
    named = dict(m.layer3d.named_parameters())
    failures = []
    _gates(_report(name, "fL", x.grad.cpu(), g["gL"], ae.grad), failures, "fL")
    _gates(_report(name, "fR", y.grad.cpu(), g["gR"], be.grad), failures, "fR")
    for k in GC_TRAIN_GRADS:
        mine, ref = golden_grad(g, k, named[k].grad)
        emu = golden_grad(g, k, pe[k].grad)[0]
        _gates(_report(name, k, mine, ref, emu), failures, k)
    assert not failures, failures
    sd = m.layer3d.state_dict()
    for mine, ref in ((sd["l33.1.running_mean"], g["rm_l33"]), (sd["l33.1.running_var"], g["rv_l33"])):
        assert float((mine.cpu() - ref).abs().max()) < 2e-2 * max(1.0, float(ref.abs().max()))


def test_frozen_batchnorm_finetuning_gradients():
    """eval-mode BatchNorm under autograd (fine-tuning with frozen statistics) runs on the fused kernels too
    (train3d.BnEvalActFunction): layer-level gradients vs torch autograd of the same fp32 graph on bf16-rounded operands"""
    import torch.nn as nn
    import torch.nn.functional as F
    from dsmnet_b200 import train3d as T
    from dsmnet_b200.volume_layout import PaddedVolume
    torch.manual_seed(4)
    B, C, D, H, W = 2, 32, 5, 6, 9
    conv = nn.Conv3d(C, C, 3, 1, 1, bias=True).cuda()
    bn = nn.BatchNorm3d(C).cuda()
    bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
    bn.eval()
    x = torch.randn(B, C, D, H, W, device="cuda").to(BF).float()
    res = torch.randn(B, C, D, H, W, device="cuda").to(BF).float()
    gz = torch.randn(B, C, D, H, W, device="cuda").to(BF).float()
    for relu in (1, 2):
        for p in list(conv.parameters()) + list(bn.parameters()):
            p.grad = None
        xv = PaddedVolume.from_ncdhw(x); xv.data.requires_grad_()
        rv = PaddedVolume.from_ncdhw(res); rv.data.requires_grad_()
        z = T.conv_bn_act(xv, conv, bn, relu, rv)
        (T.interior(z).float() * gz.permute(0, 2, 3, 4, 1)).sum().backward()
        mine = dict(w=conv.weight.grad.clone(), b=conv.bias.grad.clone(), gamma=bn.weight.grad.clone(), beta=bn.bias.grad.clone(),
                    x=PaddedVolume(xv.data.grad, B, C, D, H, W).to_ncdhw(), r=PaddedVolume(rv.data.grad, B, C, D, H, W).to_ncdhw())
        zmine = z.to_ncdhw()
        # fp32 reference on the same rounded operands
        xr = x.clone().requires_grad_(); rr = res.clone().requires_grad_()
        w = conv.weight.detach().to(BF).float().requires_grad_(); b = conv.bias.detach().clone().requires_grad_()
        ga = bn.weight.detach().clone().requires_grad_(); be = bn.bias.detach().clone().requires_grad_()
        y = F.conv3d(xr, w, None, 1, 1)
        y = y + (y.to(BF).float() - y).detach()                       # the conv kernel stores bf16 (straight-through)
        t = F.batch_norm(y + b.view(1, -1, 1, 1, 1), bn.running_mean, bn.running_var, ga, be, False, 0.1, bn.eps)
        t = (F.relu(t) + rr) if relu == 2 else F.relu(t + rr)
        (t * gz).sum().backward()
        assert float((zmine - t.detach()).abs().max()) <= 2.0 ** -6 * float(t.abs().max())
        for name, a, r in (("w", mine["w"], w.grad), ("b", mine["b"], b.grad), ("gamma", mine["gamma"], ga.grad),
                           ("beta", mine["beta"], be.grad), ("x", mine["x"], xr.grad), ("r", mine["r"], rr.grad)):
            c = cosine(a, r); ratio = float(a.norm() / r.norm())
            print("frozen-BN relu=%d grad %-5s cos %.5f ratio %.4f" % (relu, name, c, ratio))
            assert c > 0.998 and abs(ratio - 1) < 0.02
