#!/usr/bin/env python
"""Calls the memory-bound ops (corr1d, soft-argmin, imwrap; forward and backward) a few times at the BASELINE shapes,
eagerly — the workload for an `ncu -k regex:...` capture of those kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.corr1d import corr1d
from dsmnet_b200.softargmin import softargmin
from dsmnet_b200.imwrap import WarpFunction
from dsmnet_b200.imwrap import grid_vectors

dev = torch.device("cuda")
torch.manual_seed(0)
for _ in range(3):
    a = torch.relu(torch.randn(1, 128, 96, 312, device=dev)).requires_grad_(); b = torch.relu(torch.randn(1, 128, 96, 312, device=dev)).requires_grad_()
    corr1d(a, b, 41, 1).backward(torch.randn(1, 41, 96, 312, device=dev))
    c = (torch.randn(1, 192, 256, 512, device=dev) * 2).requires_grad_()
    softargmin(c, -1.0).backward(torch.randn(1, 256, 512, device=dev))
    src = torch.rand(1, 32, 540, 960, device=dev, requires_grad=True); disp = (torch.rand(1, 1, 540, 960, device=dev) * 96).requires_grad_()
    row, col = grid_vectors(540, 960, 540, 960)
    WarpFunction.apply(src, disp, row.to(dev), col.to(dev), 5e-5, False).backward(torch.randn(1, 32, 540, 960, device=dev))
torch.cuda.synchronize()
print("ok")
