#!/usr/bin/env python
"""Three eager passes of the whole drop-in PSMNet (2-D trunk on the library's kernels + hot path) at 384x1248 — the command
behind the ncu launch list / captures of the trunk kernels (profiles/r02*_whole_model_*)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda")
m = bench.synthetic_psmnet(dev)
L = torch.rand(1, 3, 384, 1248, device=dev); R = torch.rand(1, 3, 384, 1248, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = m(L, R, "test")[1]
torch.cuda.synchronize()
print("ok", float(out[0].mean()))
