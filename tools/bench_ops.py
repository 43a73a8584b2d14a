#!/usr/bin/env python
"""Per-op roofline measurements at the BASELINE.json config shapes.

    python tools/bench_ops.py [--json out.json]

Every op is called through its public wrapper (=> the C-ABI), timed with CUDA events on the
launching stream after warm-up, inputs larger than or rotated through more than L2 where it
matters.  Memory-bound ops report ALGORITHMIC bytes / time against the measured HBM copy
bandwidth (MEASURED_PEAKS.json); the 3-D convolutions report algorithmic FLOP/s against the
measured bf16 peak.  Results are copied into profiles/ by hand after each round."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def timeit(fn, reps=20, warm=3, rotate=1):
    """Device time per call of fn(i).  `rotate` consecutive calls (i = 0..rotate-1, so that a caller can
    cycle through input sets larger than L2) are captured into ONE CUDA graph and the graph is replayed:
    these kernels run for 5-100 us, far less than the Python/ctypes cost of issuing them one by one."""
    for i in range(max(warm, rotate)):
        fn(i % rotate)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(rotate):
            fn(i)
    for _ in range(warm):
        graph.replay()
    n = max(1, reps // rotate)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * rotate) * 1e-3      # seconds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from dsmnet_b200 import _lib
    from dsmnet_b200.corr1d import corr1d
    from dsmnet_b200.cost_volume import concat_volume
    from dsmnet_b200.softargmin import softargmin, upsample_softargmin
    from dsmnet_b200.imwrap import WarpFunction
    from dsmnet_b200.conv3d import FusedConv3d
    from dsmnet_b200.volume_layout import PaddedVolume
    from dsmnet_b200.imwrap import grid_vectors
    _lib.lib()
    hbm, tf, which = peaks()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    rows = []

    def mem(name, nbytes, t):
        gbs = nbytes / t / 1e9
        rows.append({"op": name, "bound": "hbm", "ms": t * 1e3, "alg_MB": nbytes / 1e6, "achieved_GBs": gbs, "frac_of_%s_hbm" % which: gbs / hbm})
        print("%-58s %8.3f ms  %9.1f MB  %8.1f GB/s  %5.1f%% of %s HBM copy peak" % (name, t * 1e3, nbytes / 1e6, gbs, 100 * gbs / hbm, which))

    def tens(name, flops, t):
        tfs = flops / t / 1e12
        rows.append({"op": name, "bound": "tensor", "ms": t * 1e3, "alg_GFLOP": flops / 1e9, "achieved_TFLOPs": tfs, "frac_of_%s_bf16" % which: tfs / tf})
        print("%-58s %8.3f ms  %9.1f GF  %8.1f TF/s  %5.1f%% of %s bf16 peak" % (name, t * 1e3, flops / 1e9, tfs, 100 * tfs / tf, which))

    # ---- op 1: corr1d (cfg1 DispNetC, cfg4 iResNet) ------------------------------------------
    for (B, C, H, W, D, s, tag) in [(1, 128, 96, 312, 41, 1, "cfg1 DispNetC"), (1, 128, 135, 240, 81, 1, "cfg4 iResNet s1"),
                                    (1, 64, 270, 480, 41, 2, "cfg4 iResNet s2"), (8, 128, 96, 312, 41, 1, "DispNetC B=8")]:
        sets = [(torch.relu(torch.randn(B, C, H, W, device=dev)), torch.relu(torch.randn(B, C, H, W, device=dev))) for _ in range(6)]
        t = timeit(lambda i: corr1d(sets[i][0], sets[i][1], D, s), rotate=6)
        mem("corr1d fwd %s (%d,%d,%d,%d) D=%d" % (tag, B, C, H, W, D), 4.0 * B * H * W * (2 * C + D), t)
        a, b = sets[0][0].clone().requires_grad_(), sets[0][1].clone().requires_grad_()
        g = torch.randn(B, D, H, W, device=dev)
        # backward = (forward + backward captured together) - forward: autograd replays on the forward's stream
        t2 = timeit(lambda i: torch.autograd.grad(corr1d(a, b, D, s), (a, b), g))
        t1 = timeit(lambda i: corr1d(a, b, D, s))
        mem("corr1d bwd %s" % tag, 4.0 * B * H * W * (4 * C + D), max(t2 - t1, 1e-9))

    # ---- op 2: concat volume ------------------------------------------------------------------
    fL = torch.randn(1, 32, 96, 312, device=dev); fR = torch.randn(1, 32, 96, 312, device=dev)
    vol = PaddedVolume.empty(1, 64, 48, 96, 312, dev, zero_rim=False)
    t = timeit(lambda i: concat_volume(fL, fR, 48, "psm", padded_bf16=True, out=vol))
    mem("concat volume PSM bf16 NDHWC (1,64,48,96,312)", 4.0 * 96 * 312 * 64 + 2.0 * 64 * 48 * 96 * 312, t)
    t = timeit(lambda i: concat_volume(fL, fR, 48, "psm"))
    mem("concat volume PSM fp32 NCDHW (1,64,48,96,312)", 4.0 * 96 * 312 * 64 + 4.0 * 64 * 48 * 96 * 312, t)
    gL = torch.randn(1, 32, 128, 256, device=dev); gR = torch.randn(1, 32, 128, 256, device=dev)
    t = timeit(lambda i: concat_volume(gL, gR, 96, "gc"))
    mem("concat volume GC fp32 NCDHW (1,64,96,128,256)", 4.0 * 128 * 256 * 64 + 4.0 * 64 * 96 * 128 * 256, t)

    # ---- op 4: soft-argmin -----------------------------------------------------------------------
    costs = [torch.randn(1, 192, 256, 512, device=dev) * 2 for _ in range(3)]
    t = timeit(lambda i: softargmin(costs[i], -1.0), rotate=3)
    mem("softargmin GC-Net (1,192,256,512)", 4.0 * 256 * 512 * 193, t)
    costs = [torch.randn(1, 192, 384, 1248, device=dev) * 2 for _ in range(2)]
    t = timeit(lambda i: softargmin(costs[i], 1.0), rotate=2)
    mem("softargmin PSMNet full-res (1,192,384,1248)", 4.0 * 384 * 1248 * 193, t)
    lr = torch.randn(3, 48, 96, 312, device=dev) * 2
    t = timeit(lambda i: upsample_softargmin(lr, (192, 384, 1248), True))
    rows.append({"op": "fused upsample+softargmin, 3 heads (3,48,96,312)->(3,384,1248)", "bound": "sfu", "ms": t * 1e3})
    print("%-58s %8.3f ms  (3 heads; moves %.1f MB; SFU/issue bound, no HBM fraction)" % ("fused upsample+softargmin x3 heads", t * 1e3, 3 * 4e-6 * (48 * 96 * 312 + 384 * 1248)))
    del costs

    # ---- op 5: imwrap ----------------------------------------------------------------------------
    for (B, C, H0, W0, tag) in [(1, 32, 540, 960, "cfg4 iResNet"), (8, 3, 384, 768, "self-sup level 0 B=8")]:
        src = torch.rand(B, C, H0, W0, device=dev); disp = torch.rand(B, 1, H0, W0, device=dev) * 0.1 * W0
        row, col = grid_vectors(H0, W0, H0, W0)
        row, col = row.to(dev), col.to(dev)
        t = timeit(lambda i: WarpFunction.apply(src, disp, row, col, 5e-5, False))
        mem("imwrap fwd %s (%d,%d,%d,%d)" % (tag, B, C, H0, W0), 4.0 * B * (2 * C * H0 * W0 + H0 * W0), t)
        s2, d2 = src.clone().requires_grad_(), disp.clone().requires_grad_()
        g = torch.randn(B, C, H0, W0, device=dev)
        t2 = timeit(lambda i: torch.autograd.grad(WarpFunction.apply(s2, d2, row, col, 5e-5, False), (s2, d2), g))
        mem("imwrap bwd %s" % tag, 4.0 * B * (4 * C * H0 * W0 + 2 * H0 * W0), max(t2 - t, 1e-9))
        # the same warp with a spatially smooth disparity field (what a network predicts): neighbouring pixels gather
        # neighbouring source pixels; the U[0, 0.1 W) field above is per-pixel random, i.e. an incoherent gather
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H0, device=dev), torch.linspace(0, 1, W0, device=dev), indexing="ij")
        smooth = ((0.05 + 0.04 * torch.sin(6.0 * xx + 3.0 * yy)) * W0).expand(B, 1, H0, W0).contiguous()
        t = timeit(lambda i: WarpFunction.apply(src, smooth, row, col, 5e-5, False))
        mem("imwrap fwd %s, smooth disparity" % tag, 4.0 * B * (2 * C * H0 * W0 + H0 * W0), t)
        d3 = smooth.clone().requires_grad_()
        t2 = timeit(lambda i: torch.autograd.grad(WarpFunction.apply(s2, d3, row, col, 5e-5, False), (s2, d3), g))
        mem("imwrap bwd %s, smooth disparity" % tag, 4.0 * B * (4 * C * H0 * W0 + 2 * H0 * W0), max(t2 - t, 1e-9))

    # ---- op 3: the PSMNet 3-D conv layer types (App. C) -------------------------------------------
    def conv_case(name, cin, cout, D, H, W, stride, transposed, res):
        w = torch.randn(cin, cout, 3, 3, 3) * 0.05 if transposed else torch.randn(cout, cin, 3, 3, 3) * 0.05
        bn = torch.nn.BatchNorm3d(cout).eval() if cout > 1 else None
        layer = FusedConv3d(w, bn, None, stride, transposed, cout > 1, dev)
        x = PaddedVolume.from_ncdhw(torch.randn(1, cin, D, H, W, device=dev))
        y = layer(x)
        r = None
        if res:
            r = torch.randn_like(y) if isinstance(y, torch.Tensor) else PaddedVolume.from_ncdhw(torch.randn(1, cout, y.D, y.H, y.W, device=dev))
        t = timeit(lambda i: layer(x, y, residual=r))
        vox = D * H * W if transposed else (y.shape[-3] * y.shape[-2] * y.shape[-1] if isinstance(y, torch.Tensor) else y.D * y.H * y.W)
        tens(name, 2.0 * 27 * cin * cout * vox, t)

    conv_case("conv 64->32 s1 @48x96x312 (dres0.0)", 64, 32, 48, 96, 312, 1, False, False)
    conv_case("conv 32->32 s1 @48x96x312 (dres0.2, dres1.0, classif.0)", 32, 32, 48, 96, 312, 1, False, False)
    conv_case("conv 32->32 s1 +residual (dres1.2)", 32, 32, 48, 96, 312, 1, False, True)
    conv_case("conv 32->64 s2 (hourglass conv1)", 32, 64, 48, 96, 312, 2, False, False)
    conv_case("conv 64->64 s1 @24x48x156 +res (conv2)", 64, 64, 24, 48, 156, 1, False, True)
    conv_case("conv 64->64 s2 (conv3)", 64, 64, 24, 48, 156, 2, False, False)
    conv_case("conv 64->64 s1 @12x24x78 (conv4)", 64, 64, 12, 24, 78, 1, False, False)
    conv_case("deconv 64->64 +res (conv5)", 64, 64, 12, 24, 78, 2, True, True)
    conv_case("deconv 64->32 +res (conv6)", 64, 32, 24, 48, 156, 2, True, True)
    conv_case("conv 32->1 s1 fp32 +res (classif.2)", 32, 1, 48, 96, 312, 1, False, True)

    # ---- whole hot paths (CUDA graph, inputs resident) ----------------------------------------------
    from dsmnet_b200.gcnet import GCNetHotPath
    from dsmnet_b200.psmnet import PSMNetHotPath
    with torch.no_grad():
        gc = GCNetHotPath(192).to(dev).eval()
        for mod in gc.modules():                     # O(1) activations with random weights: unit-variance-ish BN
            if isinstance(mod, torch.nn.BatchNorm3d):
                mod.running_var.fill_(2.0)
        gl = torch.randn(1, 32, 128, 256, device=dev); gr = torch.randn(1, 32, 128, 256, device=dev)
        t = timeit(lambda i: gc(gl, gr), reps=10)
        tens("GC-Net hot path fwd 256x512 maxdisp 192 (volume + 19 convs + head)", 882.6e9, t)
        psm = PSMNetHotPath(192).to(dev).eval()
        pl = torch.randn(1, 32, 96, 312, device=dev); pr = torch.randn(1, 32, 96, 312, device=dev)
        t = timeit(lambda i: psm(pl, pr, (384, 1248)), reps=10)
        tens("PSMNet hot path fwd 384x1248 maxdisp 192 (volume + 28 convs + 3 heads)", 926.7e9, t)

    # ---- training steps (eager autograd; convolutions fwd/dgrad/wgrad on the sm_100a kernels) ------------------
    def eager_time(fn, reps=5):
        for _ in range(3):                       # the caching allocator settles (cudaMalloc calls) over the first steps
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    gc.train()
    gtg = torch.rand(1, 1, 256, 512, device=dev) * 96

    def gc_step():
        gc.zero_grad(set_to_none=True)
        (gc(gl, gr) - gtg).abs().mean().backward()
    t = eager_time(gc_step)
    tens("GC-Net 3-D path TRAIN step fwd+bwd 256x512 (cfg3; flops = 3x fwd)", 3 * 882.6e9, t)
    print("    peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
    del gc
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    psm.train()
    gtp = torch.rand(1, 384, 1248, device=dev) * 96

    def psm_step():
        psm.zero_grad(set_to_none=True)
        sum((p - gtp).abs().mean() for p in psm(pl, pr, (384, 1248))).backward()
    t = eager_time(psm_step)
    tens("PSMNet 3-D path TRAIN step fwd+bwd 384x1248 (flops = 3x fwd)", 3 * 926.7e9, t)
    print("    peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))

    if args.json:
        json.dump({"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": which}, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
