#!/usr/bin/env python
"""Top CUDA kernels of one GC-Net 3-D-stack training step at 256x512, maxdisp 192 (torch.profiler)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.gcnet import GCNetHotPath
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
gc = GCNetHotPath(192).to(dev).train()
gl = torch.randn(1, 32, 128, 256, device=dev); gr = torch.randn(1, 32, 128, 256, device=dev)
gt = torch.rand(1, 1, 256, 512, device=dev) * 96
def step():
    gc.zero_grad(set_to_none=True)
    (gc(gl, gr) - gt).abs().mean().backward()
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=80))
