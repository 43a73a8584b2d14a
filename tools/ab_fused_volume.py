#!/usr/bin/env python
"""A/B of the fused first layer (dres0.0 builds its volume tiles itself) against the materialised volume: the hot-path
step as one CUDA graph, both ways, same process (DSM_FUSED_VOLUME is read at import, so the unfused plan is built by
flipping the module flag)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import dsmnet_b200.psmnet as P

dev = torch.device("cuda")
fL = torch.randn(1, 32, 96, 312, device=dev); fR = torch.randn(1, 32, 96, 312, device=dev)
res = {}
for fused in (True, False, True, False):
    P.FUSED_VOLUME = fused
    m = bench.synthetic_hotpath(dev)
    with torch.no_grad():
        run, graph, _ = bench.capture(lambda: m(fL, fR, (384, 1248)))
    ms = bench.time_kernel_alone(run, reps=200)
    res.setdefault(fused, []).append(ms)
    print("fused volume %-5s: %.4f ms/step (%.1f pairs/s)" % (fused, ms, 1e3 / ms))
    del m, run, graph
a, b = min(res[True]), min(res[False])
print("fused %.4f ms vs materialised %.4f ms: %.2f %% faster" % (a, b, 100 * (b - a) / b))

# the whole model from images: the trunk's last layer writes the zero-rimmed bf16 NHWC maps directly (no pack, no volume)
L = torch.rand(1, 3, 384, 1248, device=dev); R = torch.rand(1, 3, 384, 1248, device=dev)
res = {}
for fused in (True, False, True, False):
    P.FUSED_VOLUME = fused
    m = bench.synthetic_psmnet(dev)
    with torch.no_grad():
        run, graph, _ = bench.capture(lambda: m(L, R, "test")[1])
    ms = bench.time_kernel_alone(run, reps=200)
    res.setdefault(fused, []).append(ms)
    print("whole model, fused volume %-5s: %.4f ms/pair (%.1f pairs/s)" % (fused, ms, 1e3 / ms))
    del m, run, graph
a, b = min(res[True]), min(res[False])
print("whole model: fused %.4f ms vs materialised %.4f ms: %.2f %% faster" % (a, b, 100 * (b - a) / b))
