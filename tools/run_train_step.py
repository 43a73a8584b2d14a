#!/usr/bin/env python
"""Two eager training steps (fwd + bwd) of the PSMNet 3-D stack at 384x1248 — the workload for an ncu capture of the
training-only kernels (tcgen05 wgrad, fused BatchNorm, 32->1 backward, fused head backward)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.psmnet import PSMNetHotPath
dev = torch.device("cuda")
psm = PSMNetHotPath(192).to(dev).train()
pl = torch.randn(1, 32, 96, 312, device=dev, requires_grad=True); pr = torch.randn(1, 32, 96, 312, device=dev, requires_grad=True)
gt = torch.rand(1, 384, 1248, device=dev) * 96
for _ in range(2):
    psm.zero_grad(set_to_none=True)
    sum((p - gt).abs().mean() for p in psm(pl, pr, (384, 1248))).backward()
torch.cuda.synchronize()
print("ok")
