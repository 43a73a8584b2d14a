#!/usr/bin/env python
"""Where a whole PSMNet forward (384x1248, maxdisp 192) spends its time: the stock-PyTorch 2-D trunk (a caller of the
hot path) against the sm_100a hot path, eager, CUDA events."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.psmnet import PSMNet

dev = torch.device("cuda")
torch.manual_seed(0)
net = PSMNet(192).to(dev).eval()
L = torch.rand(1, 3, 384, 1248, device=dev); R = torch.rand(1, 3, 384, 1248, device=dev)


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


with torch.no_grad():
    whole = t(lambda: net(L, R, "test"))
    trunk = t(lambda: (net.feature_extraction(L), net.feature_extraction(R)))
    fL, fR = net.feature_extraction(L), net.feature_extraction(R)
    hot = t(lambda: super(PSMNet, net).forward(fL, fR, (384, 1248)))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        trunk_bf16 = t(lambda: (net.feature_extraction(L), net.feature_extraction(R)))
    fe_cl = net.feature_extraction.to(memory_format=torch.channels_last)      # only the 2-D trunk has rank-4 weights
    Lc, Rc = L.contiguous(memory_format=torch.channels_last), R.contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        trunk_bf16_cl = t(lambda: (fe_cl(Lc), fe_cl(Rc)))
print("whole PSMNet forward (eager)        : %.2f ms" % whole)
print("  2-D trunk, both images (fp32/TF32) : %.2f ms" % trunk)
print("  hot path (volume + 3-D stack + heads, eager launches): %.2f ms" % hot)
print("  2-D trunk under bf16 autocast      : %.2f ms;  + channels_last: %.2f ms" % (trunk_bf16, trunk_bf16_cl))
