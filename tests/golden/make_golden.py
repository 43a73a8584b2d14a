"""Freeze golden vectors from the reference's OWN code (run in the build container only).

    python tests/golden/make_golden.py        # needs /root/reference (or $DSMNET_REFERENCE)

Every array written here is an output of the unmodified reference executed through
oracle/refshim.py on seeded synthetic inputs; tests/test_oracle_golden.py pins oracle/ops.py to
them on CPU and tests/test_gpu_golden.py pins the CUDA kernels to them on the B200.
The reference has no golden vectors of its own, hence this file.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle.ops as O          # noqa: E402
import oracle.refshim as R      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote %-28s %6.1f KB" % (name + ".npz", os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024))


def params_digest(params):
    h = hashlib.sha256()
    for k in sorted(params):          # RNG-derived tensors only (BN statistics depend on the host's conv kernels)
        if params[k].dim() == 5:
            h.update(k.encode()); h.update(params[k].detach().numpy().tobytes())
    return h.hexdigest()


def main():
    assert R.available(), "reference tree not found"
    torch.manual_seed(1234)

    # ---- op 1: Corr1d, the three call-site configurations scaled down (+ D > W edge case) -------
    for name, (C, H, W, D, s, k) in {
        "corr1d_dispnetc": (16, 6, 40, 11, 1, 1),     # dispnetcorr.py:27  Corr1d(1, 1, 41)
        "corr1d_iresnet2": (8, 7, 36, 9, 2, 3),       # iresnet.py:69      Corr1d(3, 2, 41)
        "corr1d_d_gt_w": (4, 3, 10, 14, 1, 1),        # util_conv.py:79    loop stops at d >= W
    }.items():
        fL = torch.relu(torch.randn(2, C, H, W)).requires_grad_()
        fR = torch.relu(torch.randn(2, C, H, W)).requires_grad_()
        out = R.corr1d(fL, fR, D, s, k)
        g = torch.randn_like(out)
        out.backward(g)
        save(name, fL=fL, fR=fR, D=D, stride=s, kernel_size=k, out=out, gout=g, gL=fL.grad, gR=fR.grad)

    # ---- op 2: volumes ------------------------------------------------------------------------
    fL = torch.randn(2, 8, 5, 12); fR = torch.randn(2, 8, 5, 12)
    save("volume_psm", fL=fL, fR=fR, maxdisp=28, out=R.psm_volume(fL, fR, 28))         # D = 28//4 = 7
    save("volume_gc", fL=fL, fR=fR, D=6, out=R.gc_volume(fL, fR, 6))
    save("volume_gc_right", fL=fL, fR=fR, D=6, out=R.gc_volume(fL, fR, 6, right=True))
    a = fL.clone().requires_grad_(); b = fR.clone().requires_grad_()
    v = R.gc_volume(a, b, 6); g = torch.randn_like(v); v.backward(g)
    save("volume_gc_bwd", gout=g, D=6, gL=a.grad, gR=b.grad)

    # ---- op 4: heads ---------------------------------------------------------------------------
    mods = R._load()
    cost_lr = torch.randn(2, 1, 4, 5, 6) * 2
    with R.pinned_torch():
        import torch.nn.functional as F
        up = F.upsample(cost_lr, [16, 20, 24], mode="trilinear")
        prob = F.softmax(torch.squeeze(up, 1), dim=1)
        pred = mods["submodule"].disparityregression(16)(prob)
    save("head_psm", cost_lr=cost_lr, size=[16, 20, 24], upsampled=up, prob=prob, pred=pred)
    x37 = torch.randn(2, 1, 12, 6, 8) * 2                                      # gcnet.py:104-111
    out = torch.nn.Softmax2d()(-x37.squeeze(1))
    out = out.permute(0, 2, 3, 1).matmul(torch.arange(0, out.shape[1]).type_as(out)).unsqueeze(1)
    save("head_gc", x37=x37, pred=out)

    # ---- op 5: imwrap --------------------------------------------------------------------------
    src = torch.rand(2, 3, 20, 30)
    cases = {
        "imwrap_plain": dict(disp=torch.rand(2, 1, 20, 30) * 4, kw=dict()),
        "imwrap_fliplr": dict(disp=torch.rand(2, 1, 20, 30) * 4, kw=dict(fliplr=True)),
        "imwrap_lefttop": dict(disp=torch.rand(2, 1, 12, 18) * 3, kw=dict(LeftTop=(4, 2))),
        "imwrap_scale2": dict(disp=torch.rand(2, 1, 8, 12) * 2, kw=dict(LeftTop=(2, 2), scale_factor=2)),
        "imwrap_intdisp": dict(disp=torch.randint(0, 6, (2, 1, 20, 30)).float(), kw=dict()),
        "imwrap_oob": dict(disp=torch.rand(2, 1, 20, 30) * 60 - 20, kw=dict()),
    }
    for i, (name, c) in enumerate(cases.items()):
        s = src.clone().requires_grad_(); d = c["disp"].clone().requires_grad_()
        out, delt = R.imwrap(s, d, seed=100 + i, **c["kw"])
        g = torch.randn_like(out)
        out.backward(g)
        kw = c["kw"]
        save(name, src=src, disp=c["disp"], fliplr=bool(kw.get("fliplr", False)), LeftTop=list(kw.get("LeftTop", (0, 0))),
             scale_factor=kw.get("scale_factor", 1), delt=delt, out=out, gout=g, gsrc=s.grad, gdisp=d.grad)

    # ---- op 3 + the north-star path: the reference's PSMNet 3-D stack ----------------------------
    maxdisp, H, W = 32, 32, 80
    rs = np.random.RandomState(2024)
    fL = torch.from_numpy(rs.standard_normal((1, 32, H // 4, W // 4)).astype(np.float32))
    fR = torch.from_numpy(rs.standard_normal((1, 32, H // 4, W // 4)).astype(np.float32))
    cost = R.psm_volume(fL, fR, maxdisp)
    params = O.psmnet_random_params(seed=7, calibrate_on=cost)
    net = R.make_psmnet(maxdisp, seed=0)
    missing = net.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all(k.startswith("feature_extraction") or k.endswith("num_batches_tracked") for k in missing.missing_keys)
    c1, c2, c3 = R.psmnet_3d(net, cost)
    p1, p2, p3 = R.psmnet_heads(net, (c1, c2, c3), maxdisp, H, W)
    save("psmnet_hotpath", fL=fL, fR=fR, maxdisp=maxdisp, H=H, W=W, seed=7, params_sha256=params_digest(params),
         cost1=c1, cost2=c2, cost3=c3, pred1=p1, pred2=p2, pred3=p3)

    # ---- GC-Net 3-D path: the reference's own feature3d (eval-mode BatchNorm) on a GC volume -----------
    maxdisp = 16                                                      # D = maxdisp/2 = 8 (gcnet.py:117)
    rs = np.random.RandomState(2025)
    gfL = torch.from_numpy(rs.standard_normal((1, 32, 16, 24)).astype(np.float32))
    gfR = torch.from_numpy(rs.standard_normal((1, 32, 16, 24)).astype(np.float32))
    gvol = R.gc_volume(gfL, gfR, maxdisp // 2)
    gparams = O.gcnet_random_params(seed=9, calibrate_on=gvol)
    gnet = R.make_gcnet(maxdisp, seed=0).eval()
    gmissing = gnet.load_state_dict({"layer3d." + k: v for k, v in gparams.items()}, strict=False)
    assert not gmissing.unexpected_keys, gmissing.unexpected_keys
    assert all(k.startswith("layer2d") or k.endswith("num_batches_tracked") for k in gmissing.missing_keys)
    with torch.no_grad(), R.pinned_torch():
        gdisp = gnet.layer3d(gvol, "test")                            # [1,1,32,48]
    save("gcnet_hotpath", fL=gfL, fR=gfR, maxdisp=maxdisp, seed=9, params_sha256=params_digest(gparams), disp=gdisp)

    # ---- DispNetC (BASELINE config 1): the reference's dispnetcorr with its Corr1d layer, 7-level pyramid ------
    dparams = O.dispnetc_random_params(seed=5)
    dnet = mods["dispnetcorr"].dispnetcorr(192).eval()
    dmissing = dnet.load_state_dict(dparams, strict=False)
    assert not dmissing.unexpected_keys and not dmissing.missing_keys, dmissing
    rs = np.random.RandomState(1)
    dimL = torch.from_numpy(rs.standard_normal((1, 3, 128, 192)).astype(np.float32))
    dimR = torch.from_numpy(rs.standard_normal((1, 3, 128, 192)).astype(np.float32))
    with torch.no_grad(), R.pinned_torch():
        _, douts = dnet(dimL, dimR, "test")
    save("dispnetc_forward", imL=dimL, imR=dimR, seed=5, **{"out%d" % i: o for i, o in enumerate(douts)})

    # ---- iResNet (BASELINE config 4): the reference's iresnet (Corr1d D=81, imwrap feature constancy, Corr1d k3 s2) -------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import random_state_dict
    from dsmnet_b200.iresnet import iresnet as _ires_shape          # only to enumerate parameter names/shapes
    isd = random_state_dict(_ires_shape(192), 11)
    inet = mods["iresnet"].iresnet(192).eval()
    imiss = inet.load_state_dict(isd, strict=True)
    rs = np.random.RandomState(2)
    iimL = torch.from_numpy(rs.standard_normal((1, 3, 128, 192)).astype(np.float32))
    iimR = torch.from_numpy(rs.standard_normal((1, 3, 128, 192)).astype(np.float32))
    with torch.no_grad(), R.pinned_torch():
        torch.manual_seed(7)                                         # the warp draws delt from the global RNG (imwrap.py:70)
        iscales, iouts = inet(iimL, iimR, "test")
    save("iresnet_forward", imL=iimL, imR=iimR, seed=11, rng_seed=7, scales=np.asarray(iscales),
         **{"out%d" % i: o for i, o in enumerate(iouts)})

    # ---- self-supervised pyramid loss (losses/loss.py `depthmono-mask`, 28 imwrap calls) run by the reference itself ----
    rs = np.random.RandomState(0)
    sB, sh, sw, ne = 2, 64, 128, 16
    u = lambda *shape: torch.from_numpy(rs.uniform(0, 1, size=shape).astype(np.float32))
    sbatch = u(sB, 6, sh + 2 * ne, sw + 2 * ne)
    sb1 = torch.flip(sbatch, dims=[3])
    crop = lambda t: t[:, :, ne:-ne, ne:-ne]
    sdisp = [u(sB, 1, sh >> l, sw >> l) * 6 / (2 ** l) for l in range(7)]
    sdisp1 = [u(sB, 1, sh >> l, sw >> l) * 6 / (2 ** l) for l in range(7)]
    wl = [1.0, 0.5, 0.3, 0.2, 0.1, 0.05, 0.01]
    a = [x.clone().requires_grad_() for x in sdisp]; a1 = [x.clone().requires_grad_() for x in sdisp1]
    sloss = R.selfsup_loss("depthmono-mask", list(range(7)), a, a1, crop(sbatch[:, :3]), sbatch[:, 3:6], crop(sb1[:, 3:6]), sb1[:, :3],
                           (ne, ne), wl, seed=3)
    sloss.backward()
    save("selfsup_loss", batch=sbatch, nedge=ne, seed=3, weight_levels=np.asarray(wl, dtype=np.float32), loss=sloss.detach(),
         **{"disp%d" % l: sdisp[l] for l in range(7)}, **{"disp1_%d" % l: sdisp1[l] for l in range(7)},
         **{"g%d" % l: a[l].grad for l in range(7)}, **{"g1_%d" % l: a1[l].grad for l in range(7)})

    # single reference layers (GC-Net style: bias + BN + ReLU; stride 2; transposed with BN3d swap)
    uc = mods["util_conv"]
    torch.manual_seed(77)
    x = torch.randn(1, 32, 5, 6, 9)
    for name, layer in {"gc_conv_s2": uc.conv3d_bn(32, 32, 3, 2, True, True, torch.nn.ReLU()),
                        "gc_deconv": uc.deconv3d_bn(32, 32, 3, 2, True, True, torch.nn.ReLU())}.items():
        for i, m in enumerate(layer):
            if isinstance(m, torch.nn.BatchNorm2d):
                layer[i] = torch.nn.BatchNorm3d(m.num_features)
        bn = layer[1]
        bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5)
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.2)
        layer.eval()
        y = layer(x)
        save(name, x=x, weight=layer[0].weight, bias=layer[0].bias, bn_weight=bn.weight, bn_bias=bn.bias,
             bn_mean=bn.running_mean, bn_var=bn.running_var, y=y)


if __name__ == "__main__":
    main()
