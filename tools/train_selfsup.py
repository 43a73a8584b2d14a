#!/usr/bin/env python
"""BASELINE.json config 5b: the self-supervised DispNetC training step (stereo_selfsupervised.py:60-100 with the
`depthmono-mask` loss: two model forwards, 28 imwrap warps, SSIM/smoothness/left-right terms, Adam) on synthetic
768x384 crops (net input 640x256 after the 64-pixel edge, DSMnet_train_kitti-raw.sh:10-11), one process per GPU,
batch sharded by rank, gradients (42.2 M fp32 = 168.7 MB) all-reduced over NCCL by DistributedDataParallel.

  python tools/train_selfsup.py --batch 4 --steps 10                       # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_selfsup.py --batch 4

Prints one JSON line (rank 0): images/s over all ranks, ms per step (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4, help="stereo pairs per GPU per step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from dsmnet_b200.dispnetcorr import dispnetcorr
    from dsmnet_b200.selfsup import train_step
    from dsmnet_b200.shard import env_rank_world
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = dispnetcorr(192).to(dev).train()
    nparams = sum(p.numel() for p in model.parameters())
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=32, gradient_as_bucket_view=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1000 + rank)
    batch = torch.rand(args.batch, 6, 384, 768, generator=g).to(dev)
    wl = [1.0, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01]                   # Weight_Adjust_levels at the end of the schedule (loss.py:372-386)
    for _ in range(args.warmup):
        loss = train_step(model, opt, batch, 64, wl)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = train_step(model, opt, batch, 64, wl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        print(json.dumps({"workload": "dispnetc_selfsup_train_step_768x384_nedge64", "n_gpus": world, "batch_per_gpu": args.batch,
                          "ms_per_step": ms, "pairs_per_s": world * args.batch / (ms / 1e3), "loss": float(loss),
                          "params": nparams, "grad_allreduce_MB": nparams * 4 / 1e6 if world > 1 else 0.0}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
