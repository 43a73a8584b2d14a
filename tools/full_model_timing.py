#!/usr/bin/env python
"""Where a whole PSMNet forward (384x1248, maxdisp 192) spends its time: the 2-D trunk on the library's kernels
(dsmnet_b200/trunk2d.py) and as stock PyTorch / cuDNN (fp32-TF32, bf16 autocast, channels_last), the hot path, the whole
model eager and as one CUDA graph; plus a per-layer table of the trunk's tensor-core convolutions.  CUDA events.

    python tools/full_model_timing.py [--layers]
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.psmnet import PSMNet
from dsmnet_b200 import trunk2d

dev = torch.device("cuda")
torch.manual_seed(0)
net = PSMNet(192).to(dev).eval()
for m in net.modules():                                   # keep activations O(1) through 56 layers
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_var.fill_(2.0)
L = torch.rand(1, 3, 384, 1248, device=dev); R = torch.rand(1, 3, 384, 1248, device=dev)
LR = torch.cat((L, R), 0)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fe = net.feature_extraction
with torch.no_grad():
    whole = t(lambda: net(L, R, "test"))
    trunk = t(lambda: fe(LR))
    fea = fe(LR)
    fL, fR = fea[:1].contiguous(), fea[1:].contiguous()
    hot = t(lambda: super(PSMNet, net).forward(fL, fR, (384, 1248)))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = net(L, R, "test")
    whole_graph = t(lambda: g.replay(), 20)
    gt = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gt):
        f2 = fe(LR)
    trunk_graph = t(lambda: gt.replay(), 20)
    with torch.enable_grad():                             # the module's stock-PyTorch graph (cuDNN)
        stock = t(lambda: (fe(L), fe(R)))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stock_bf16 = t(lambda: (fe(L), fe(R)))
        fe_cl = fe.to(memory_format=torch.channels_last)
        Lc, Rc = L.contiguous(memory_format=torch.channels_last), R.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stock_bf16_cl = t(lambda: (fe_cl(Lc), fe_cl(Rc)))
flops = 424e9
print("whole PSMNet forward, eager          : %.3f ms   as one CUDA graph: %.3f ms (%.0f pairs/s)" % (whole, whole_graph, 1e3 / whole_graph))
print("  2-D trunk on own kernels, both images as one batch: %.3f ms eager, %.3f ms graph (%.0f TFLOP/s of 424 GFLOP)" %
      (trunk, trunk_graph, flops / trunk_graph / 1e9))
print("  hot path (volume + 3-D stack + heads), eager      : %.3f ms" % hot)
print("  stock PyTorch trunk (cuDNN): fp32/TF32 %.2f ms; bf16 autocast %.2f ms; + channels_last %.2f ms" % (stock, stock_bf16, stock_bf16_cl))

if "--layers" in sys.argv:
    plan = trunk2d.cached_plan(fe, trunk2d.PSMNetTrunkPlan, dev)
    ws = plan._workspace(2, 384, 1248)
    half, q64, q128, cat = ws["half"], ws["q64"], ws["q128"], ws["cat"]
    rows = []
    def layer(name, fn, fl, conv=None):
        """20 back-to-back launches captured in one CUDA graph (the way the plan runs them)"""
        def once():
            gl = torch.cuda.CUDAGraph()
            fn(); torch.cuda.synchronize()
            with torch.cuda.graph(gl):
                for _ in range(20):
                    fn()
            return t(lambda: gl.replay(), 10) / 20
        ms = once()
        rows.append((name, ms * 1e3, fl / ms / 1e9))
        if conv is not None and conv.k == 3 and conv.stride == 1 and conv.cin <= 128 and conv.cout <= 128:
            v0 = conv.variant
            for tag, v in (("  .. per-tile kernel (variant 4)", v0 | 4), ("  .. row-sharing, one MMA issuer (variant 8)", v0 | 8)):
                conv.variant = v
                ms = once()
                rows.append((tag, ms * 1e3, fl / ms / 1e9))
            conv.variant = v0
    px2, px4 = 2 * 192 * 624, 2 * 96 * 312
    layer("firstconv.0 3->32 s2 (CUDA cores)", lambda: plan.first(LR, half[0]), 2 * 27 * 32 * px2)
    layer("32->32 k3 @192x624", lambda: plan.fc2(half[0], half[1]), 2 * 9 * 32 * 32 * px2, plan.fc2)
    c1, c2, ds = plan.layers[1][0]
    layer("32->64 k3 s2", lambda: c1(half[0], q64[0]), 2 * 9 * 32 * 64 * px4)
    layer("32->64 k1 s2", lambda: ds(half[0], q64[1]), 2 * 32 * 64 * px4)
    c1, c2, ds = plan.layers[1][1]
    layer("64->64 k3 @96x312", lambda: c1(q64[0], q64[1]), 2 * 9 * 64 * 64 * px4, c1)
    layer("64->64 k3 + residual", lambda: c2(q64[0], q64[1], residual=q64[2]), 2 * 9 * 64 * 64 * px4, c2)
    c1, c2, ds = plan.layers[2][0]
    layer("64->128 k3 (from cat slice)", lambda: c1(cat, q128[0]), 2 * 9 * 64 * 128 * px4, c1)
    layer("64->128 k1", lambda: ds(cat, q128[1]), 2 * 64 * 128 * px4)
    c1, c2, ds = plan.layers[2][1]
    layer("128->128 k3", lambda: c1(q128[0], q128[1]), 2 * 9 * 128 * 128 * px4, c1)
    c1, c2, ds = plan.layers[3][1]
    layer("128->128 k3 dilation 2", lambda: c1(q128[0], q128[1]), 2 * 9 * 128 * 128 * px4, c1)
    layer("320->128 k3 (lastconv.0)", lambda: plan.last0(cat, q128[0]), 2 * 9 * 320 * 128 * px4)
    o = torch.empty(2, 32, 96, 312, device=dev)
    layer("128->32 k1 -> fp32 NCHW", lambda: plan.last2(q128[0], o), 2 * 128 * 32 * px4)
    print("%-52s %10s %10s" % ("trunk layer (batch 2)", "us", "TFLOP/s"))
    for name, us, tf in rows:
        print("%-52s %10.1f %10.1f" % (name, us, tf))
