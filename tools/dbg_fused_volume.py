#!/usr/bin/env python
"""Which stage bounds the fused first layer?  Times dsm_conv3d_volume_fwd at the BASELINE size with parts of the kernel switched
off through the `variant` debug bits (4: no epilogue traffic, 5: no global loads in the builder warps, 6: no MMAs), next to the
materialised route (concat kernel + plane-sharing conv)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.conv3d import FusedConv3d, conv_from_features, pack_features_nhwc
from dsmnet_b200.cost_volume import concat_volume
from dsmnet_b200.volume_layout import PaddedVolume

dev = torch.device("cuda")
torch.manual_seed(0)
fL = torch.randn(1, 32, 96, 312, device=dev); fR = torch.randn(1, 32, 96, 312, device=dev)
w = torch.randn(32, 64, 3, 3, 3, device=dev) * 0.03
hL, hR = pack_features_nhwc(fL), pack_features_nhwc(fR)
out = PaddedVolume.empty(1, 32, 48, 96, 312, dev)
vol = PaddedVolume.empty(1, 64, 48, 96, 312, dev, zero_rim=False)


def t(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, dbg in (("full", 0), ("no epilogue traffic", 1), ("no builder loads", 2), ("no MMAs", 4), ("builders only (no MMA, no epilogue)", 5),
                  ("bare pipeline", 7)):
    layer = FusedConv3d(w, None, None, 1, False, 1, variant=dbg << 4)
    print("fused  %-40s %8.1f us" % (name, t(lambda: conv_from_features(layer, hL, hR, 48, "psm", out))))
layer = FusedConv3d(w, None, None, 1, False, 1)
print("materialised: concat %.1f us + conv %.1f us; pack x2 %.1f us" % (
    t(lambda: concat_volume(fL, fR, 48, "psm", padded_bf16=True, out=vol)), t(lambda: layer(vol, out)),
    t(lambda: (pack_features_nhwc(fL), pack_features_nhwc(fR)))))
