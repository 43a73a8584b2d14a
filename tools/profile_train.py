#!/usr/bin/env python
"""Top CUDA kernels of one PSMNet 3-D-stack training step at 384x1248 (torch.profiler)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.psmnet import PSMNetHotPath
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
psm = PSMNetHotPath(192).to(dev).train()
pl = torch.randn(1, 32, 96, 312, device=dev); pr = torch.randn(1, 32, 96, 312, device=dev)
gt = torch.rand(1, 384, 1248, device=dev) * 96
def step():
    psm.zero_grad(set_to_none=True)
    sum((p - gt).abs().mean() for p in psm(pl, pr, (384, 1248))).backward()
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
