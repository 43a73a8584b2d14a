#!/usr/bin/env python
"""Benchmark of the north-star path: PSMNet (maxdisp 192) cost-volume hot path at 384x1248.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B]

A "step" is one pass of the hot path (concat volume -> stacked-hourglass 3-D convs -> three
upsample+soft-argmin heads; reference models/psmnet/stackhourglass.py:123-168) over a batch of
B synthetic stereo pairs per GPU (default 1: BASELINE.json configs[1]).  Prints ONE JSON line
(rank 0).  Keys follow the driver's contract:
  value     whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the public API with HOST (pinned) feature maps: H2D of both
            feature maps + the path + D2H of the three disparity maps inside the timed region
  roofline  the dominant kernel (the 32->32 k3 tcgen05 convolution): algorithmic FLOP/s measured
            live with CUDA events on the launching stream vs the measured bf16 peak
  cpu_baseline  the oracle (CPU restatement of the reference path) timed on this host's cores on
            a bounded sample (rank 0, N=1 only)
`--impl reference` times that CPU path alone (same metric/config) — the reference is pure
Python/PyTorch and cannot travel to the GPU box, so the arm runs the oracle port
(oracle/ops.py); kind = "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stereo pairs/sec at 384x1248 (PSMNet D=192) hot path: cost volume + stacked-hourglass 3-D convs + soft-argmin"
H_IMG, W_IMG, MAXDISP, C_FEAT = 384, 1248, 192, 32
SAMPLE_ROWS = 384         # CPU sample: one whole pair (all 384 image rows; ~3-4 s per pass on 16 host cores)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms (a timed region of a few
    tens of milliseconds still yields a median), nvidia-smi every 100 ms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    # NVML clocks-event-reason bits (nvml.h)
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.active = index, [], False, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            if self.nvml is not None:
                try:
                    mhz = self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
                    try:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    if self.active:
                        self.samples.append((mhz, self.max_mhz, mask))
                except Exception:
                    pass
                time.sleep(0.002)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out and self.active:
                    f = [x.strip() for x in out.split(",")]
                    mask = sum(bit for name, bit in self.BITS.items()
                               if f[2 + ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"].index(name)].lower().startswith("active"))
                    self.samples.append((int(f[0]), int(f[1]), mask))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = sorted({n for s in self.samples for n, bit in self.BITS.items() if s[2] & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synthetic_hotpath(device):
    """Random-init 3-D stack of the reference architecture (stackhourglass.py:73-114 init); the
    BatchNorm running variance is set to the analytic output variance of its He-initialised conv
    (2*Cin/Cout) so that eval-mode activations stay O(1)."""
    from dsmnet_b200.psmnet import PSMNetHotPath
    torch.manual_seed(0)
    m = PSMNetHotPath(MAXDISP)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Sequential) and len(mod) >= 2 and isinstance(mod[1], torch.nn.BatchNorm3d):
            w = mod[0].weight
            cin, cout = (w.shape[0], w.shape[1]) if isinstance(mod[0], torch.nn.ConvTranspose3d) else (w.shape[1], w.shape[0])
            if isinstance(mod[0], torch.nn.ConvTranspose3d):
                mod[0].weight.data.normal_(0, (2.0 / (27 * cout)) ** 0.5)
                mod[1].running_var.fill_(cin / (4.0 * cout))      # 27/8 taps reach an output voxel on average
            else:
                mod[1].running_var.fill_(2.0 * cin / cout)
    return m.to(device).eval()


def cpu_sample_pairs_per_s(threads=None, reps=1):
    """The oracle port on host cores: hot path on the top SAMPLE_ROWS rows of one pair."""
    import oracle.ops as O
    torch.set_num_threads(threads or os.cpu_count())
    g = torch.Generator().manual_seed(0)
    h, w = SAMPLE_ROWS // 4, W_IMG // 4
    fL = torch.randn(1, C_FEAT, h, w, generator=g); fR = torch.randn(1, C_FEAT, h, w, generator=g)
    params = O.psmnet_random_params(seed=0)      # timing does not depend on the values
    best = None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            O.psmnet_hotpath(params, fL, fR, MAXDISP, (SAMPLE_ROWS, W_IMG))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    frac = SAMPLE_ROWS / H_IMG
    return frac / best, best, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    times = []
    nthreads = os.cpu_count()
    for i in range(args.warmup + args.steps):
        v, dt, nthreads = cpu_sample_pairs_per_s()
        if i >= args.warmup:
            times.append(dt)
    frac = SAMPLE_ROWS / H_IMG
    ms = 1e3 * sum(times) / len(times)
    value = frac / (ms / 1e3)
    sample = "one full %dx%d pair per step through the oracle port of the hot path, fp32, torch CPU ops, %d threads" % (H_IMG, W_IMG, nthreads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "psmnet_hotpath_384x1248_maxdisp192", "batch_per_gpu": args.batch, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_kernel_alone(fn, reps=20):
    for _ in range(3):
        fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_native(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)
    from dsmnet_b200 import _lib
    from dsmnet_b200.conv3d import conv_timeouts
    _lib.lib()                                  # fail loudly if the CUDA library is missing
    m = synthetic_hotpath(device)
    B, h, w = args.batch, H_IMG // 4, W_IMG // 4
    g = torch.Generator().manual_seed(1000 + rank)
    host_L = torch.randn(B, C_FEAT, h, w, generator=g).pin_memory()
    host_R = torch.randn(B, C_FEAT, h, w, generator=g).pin_memory()
    fL = host_L.to(device); fR = host_R.to(device)
    host_out = [torch.empty(B, H_IMG, W_IMG).pin_memory() for _ in range(3)]
    launches_per_step = 1 + 28 + 3              # concat volume, 28 conv blocks, three head launches (each as soon as its cost exists)

    def step():
        with torch.no_grad():
            return m(fL, fR, (H_IMG, W_IMG))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"): the step captured once in a CUDA graph ------------
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):    # NCCL's watchdog thread may touch CUDA meanwhile
                static_out = step()
        except Exception as e:                  # eager launches are still the same kernels
            graph = None
            sys.stderr.write("CUDA graph capture unavailable (%s); timing eager launches\n" % e)
    run = (lambda: graph.replay()) if graph is not None else step
    for _ in range(max(args.warmup, 3)):
        run()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sampler.active = True
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    sampler.active = False
    t_ms = e0.elapsed_time(e1)

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    # Every step copies that step's two feature maps from pinned host memory and reads its three disparity maps back
    # into pinned host memory.  The copies are double-buffered on their own streams so that the H2D of step i+1 and the
    # D2H of step i-1 overlap the compute of step i (throughput metric; the host waits for the results of step i-1
    # before it enqueues step i+1, i.e. it does read every result).
    s_in, s_cmp, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device), torch.cuda.Stream(device)
    stage_in = [(torch.empty_like(fL), torch.empty_like(fR)) for _ in range(2)]
    stage_out = [[torch.empty(B, H_IMG, W_IMG, device=device) for _ in range(3)] for _ in range(2)]
    host_outs = [[torch.empty(B, H_IMG, W_IMG).pin_memory() for _ in range(3)] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]; ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]; ev_done = [torch.cuda.Event() for _ in range(2)]
    torch.cuda.synchronize()

    def e2e_enqueue(i):
        k = i & 1
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_used[k])                       # step i-2 has consumed this staging pair
            stage_in[k][0].copy_(host_L, non_blocking=True); stage_in[k][1].copy_(host_R, non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            fL.copy_(stage_in[k][0], non_blocking=True); fR.copy_(stage_in[k][1], non_blocking=True)
            ev_used[k].record(s_cmp)
            if graph is not None:
                graph.replay(); preds = static_out
            else:
                preds = step()
            s_cmp.wait_event(ev_done[k])                      # step i-2's results have left this staging triple
            for dst, src in zip(stage_out[k], preds):
                dst.copy_(src, non_blocking=True)
            ev_out[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_out[k])
            for dst, src in zip(host_outs[k], stage_out[k]):
                dst.copy_(src, non_blocking=True)
            ev_done[k].record(s_out)

    def e2e_run(n):
        for i in range(n):
            e2e_enqueue(i)
            if i >= 1:
                ev_done[(i - 1) & 1].synchronize()            # the caller reads the result of step i-1
        ev_done[(n - 1) & 1].synchronize()

    e2e_run(4)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    t_e2e = time.perf_counter() - t0
    sampler.active = False
    sampler.stop_flag = True; sampler.join(timeout=2)

    if dist is not None:
        tt = torch.tensor([t_ms, t_e2e * 1e3], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, t_e2e = float(tt[0]), float(tt[1]) / 1e3
    timeouts = conv_timeouts()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (7 of the 28 conv launches: Conv3d 32->32 @48x96x312) ----
    hbm, tf_burst, tf_sust, which = measured_peaks()
    plan = m._get_plan(device)
    ws = m._workspace(B, MAXDISP // 4, h, w, device)
    layer = plan.dres0_2
    ms_k = time_kernel_alone(lambda: layer(ws["a"], ws["c0"]))
    flops = 2.0 * 27 * 32 * 32 * B * (MAXDISP // 4) * h * w
    achieved = flops / (ms_k * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "conv3d_rs_kernel<KC=32,NP=32> (Conv3d 32->32 k3 s1 + BN + ReLU @48x96x312, 6 launches/step)",
                "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                "peak_source": which + " bf16 burst (kernel timed alone)", "ms_per_launch": ms_k,
                "traffic": 144.7e6, "traffic_source": "dram read+write of this kernel per launch, ncu --set full (profiles/r01h_prof_conv3d_rs_bench_summary.txt); algorithmic 184 MB, part of the output is still in L2 at kernel end"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, nthreads = cpu_sample_pairs_per_s(reps=3)
        cpu = {"value": v, "unit": "pairs/s", "cores": nthreads, "kind": "port",
               "sample": "oracle port (fp32 torch CPU ops) of the whole hot path on one full %dx%d pair, best of 3 passes of %.1f s" %
                         (H_IMG, W_IMG, dt)}

    pairs = B * world * args.steps
    def nbytes(v):
        t = v if isinstance(v, torch.Tensor) else v.data
        return t.numel() * t.element_size()
    act_bytes = sum(nbytes(v) for k, v in ws.items() if not isinstance(v, list)) + \
        sum(nbytes(x) for k in ("out", "pre", "post") for x in ws[k])
    line = {"metric": METRIC, "value": pairs / (t_ms / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "psmnet_hotpath_384x1248_maxdisp192", "batch_per_gpu": B,
                       "volume": "64x48x96x312 per pair", "sharding": "stereo pairs over ranks, no data-path collective",
                       "l2": "per-step activation working set %.2f GB >> 126 MB L2 (no explicit flush)" % (act_bytes / 1e9),
                       "cuda_graph": graph is not None, "conv_timeouts": timeouts},
            "clocks": sampler.summary(),
            "e2e": {"value": pairs / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * host_L.numel() * 4,
                    "d2h_bytes_per_step": 3 * host_out[0].numel() * 4},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="stereo pairs per GPU per step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
