#!/usr/bin/env python
"""Wall-clock and device time of one training step (fwd + bwd) of the PSMNet / GC-Net 3-D stacks:
eager (Python issues every launch) and replayed from one CUDA graph of the whole step; optional host profile."""
import argparse, cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsmnet_b200.psmnet import PSMNetHotPath
from dsmnet_b200.gcnet import GCNetHotPath


def measure(name, step, reps=5):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    print("%-28s eager : %.2f ms/step wall, %.2f ms/step between events" % (name, wall, e0.elapsed_time(e1) / reps))
    # whole step as one CUDA graph (static inputs, .grad tensors live in the graph's pool)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        step()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    print("%-28s graph : %.2f ms/step wall, %.2f ms/step between events" % (name, wall, e0.elapsed_time(e1) / reps))
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host-profile", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda")
    psm = PSMNetHotPath(192).to(dev).train()
    pl = torch.randn(1, 32, 96, 312, device=dev, requires_grad=True); pr = torch.randn(1, 32, 96, 312, device=dev, requires_grad=True)
    gt = torch.rand(1, 384, 1248, device=dev) * 96

    def psm_step():
        psm.zero_grad(set_to_none=True); pl.grad = pr.grad = None
        sum((p - gt).abs().mean() for p in psm(pl, pr, (384, 1248))).backward()

    if args.host_profile:
        psm_step(); torch.cuda.synchronize()
        pr_ = cProfile.Profile(); pr_.enable()
        for _ in range(3):
            psm_step()
        torch.cuda.synchronize(); pr_.disable()
        pstats.Stats(pr_).sort_stats("cumulative").print_stats(35)
    measure("PSMNet 384x1248 D=192", psm_step)
    del psm
    torch.cuda.empty_cache()
    gc = GCNetHotPath(192).to(dev).train()
    gl = torch.randn(1, 32, 128, 256, device=dev, requires_grad=True); gr = torch.randn(1, 32, 128, 256, device=dev, requires_grad=True)
    gtg = torch.rand(1, 1, 256, 512, device=dev) * 96

    def gc_step():
        gc.zero_grad(set_to_none=True); gl.grad = gr.grad = None
        (gc(gl, gr) - gtg).abs().mean().backward()
    measure("GC-Net 256x512 D=192", gc_step)


if __name__ == "__main__":
    main()
