#!/usr/bin/env python
"""Throughput sweep of BASELINE configs 5a / 5b on the N GPUs of this box: runs bench.py workloads one
after the other — under torchrun for N > 1, one rank per GPU — and appends their JSON lines to gpurun_out/r02_sweep_N<N>.jsonl.

    python tools/sweep.py N            # N = number of GPUs visible (1, 2, 4, 8)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
out = os.path.join(ROOT, "gpurun_out", "r02_sweep_N%d.jsonl" % N)
os.makedirs(os.path.dirname(out), exist_ok=True)
light = ["--no-cpu-baseline", "--no-torch-baseline", "--no-whole-model"]
runs = [
    ["--workload", "psmnet_384x1248", "--batch", "1"] + light,
    ["--workload", "psmnet_384x1248", "--batch", "8"] + light,
    ["--workload", "psmnet_540x960", "--batch", "1"] + light,
    ["--workload", "psmnet_540x960", "--batch", "8"] + light,
    ["--workload", "psmnet_full_384x1248", "--batch", "1"] + light,
    ["--workload", "psmnet_full_384x1248", "--batch", "8"] + light,
    ["--workload", "dispnetc_selfsup_train", "--batch", "4", "--steps", "10"],
    ["--workload", "gcnet_256x512", "--batch", "1"],
]
if N == 1:
    runs += [["--workload", "psmnet_384x1248", "--batch", str(b)] + light for b in (16, 32, 64)]
port = 29500
for r in runs:
    port += 1
    if N > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(N), "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(N)] + r
    else:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1"] + r
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        rec = json.loads(lines[-1]) if lines else {"error": (p.stderr or p.stdout)[-600:]}
    except Exception as e:
        rec = {"error": str(e)[:300]}
    rec["_args"] = r
    with open(out, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print(N, r[:4], rec.get("value"), rec.get("unit"), "ms/step", rec.get("ms_per_step"), rec.get("error", "")[:200], flush=True)
