#!/usr/bin/env python
"""Benchmark of the stereo cost-volume hot path and its callers on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME] [--batch B]

Workloads (one JSON line each, rank 0):
  psmnet_384x1248        (default; BASELINE.json configs[1], the metric's own configuration) — one pass of the PSMNet hot
                         path: concat volume -> stacked-hourglass 3-D convs -> three upsample+soft-argmin heads
                         (reference models/psmnet/stackhourglass.py:123-168) on B stereo pairs per GPU
  psmnet_540x960         the same path at the SceneFlow shape (odd 135x240 feature maps: the crop-to-min adds)
  psmnet_full_384x1248   the whole drop-in model from IMAGES: 2-D trunk (both images as one batch) + the hot path
  gcnet_256x512          BASELINE configs[2]: GC-Net 3-D path forward + backward (train-mode BatchNorm) on one crop
  iresnet_ops_540x960    BASELINE configs[3]: Corr1d D=81 + imwrap feature warp + Corr1d (k3, s2) at iResNet's shapes
  dispnetc_selfsup_train BASELINE configs[4b]: self-supervised DispNetC step (imwrap + SSIM loss, Adam), NCCL gradient
                         all-reduce (DistributedDataParallel) when N > 1

Keys follow the driver's contract:
  value     whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks (a short burst of K steps)
  sustained the same metric over >= 2 s of back-to-back steps with the SM clock sampled alongside
  e2e       the same metric through the package's host-buffer API (dsmnet_b200.pipeline.HostPipeline via
            PSMNetHotPath.host_pipeline / PSMNet.image_pipeline): H2D of the step's inputs from pinned host memory, the
            path, D2H of its results, all inside the timed region
  roofline  the dominant kernel timed alone with CUDA events vs the measured bf16 peak; `traffic` is read from the ncu
            capture summarised in profiles/ (tools/ncu_traffic.py), not typed in
  cpu_baseline   the oracle (CPU restatement of the reference path) on this host's cores, bounded sample (rank 0, N=1)
  torch_gpu_baseline   stated context, default workload only: the SAME path as stock PyTorch/cuDNN modules on this GPU
            (fp32 without TF32, and bf16 autocast + channels_last_3d)
`--impl reference` times the CPU path alone (same metric/config): the reference is pure Python/PyTorch and cannot travel to
the GPU box, so the arm runs the oracle port (oracle/ops.py); kind = "port".
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
MAXDISP, C_FEAT = 192, 32
WORKLOADS = ("psmnet_384x1248", "psmnet_540x960", "psmnet_full_384x1248", "gcnet_256x512", "iresnet_ops_540x960",
             "dispnetc_selfsup_train")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def profiled_traffic(kernel_substr):
    """DRAM read+write bytes per launch of a kernel, from the ncu capture summarised by tools/ncu_traffic.py into
    profiles/r02_roofline_traffic.json (regenerated whenever the kernel changes); None when there is no capture."""
    p = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    if not os.path.isfile(p):
        return None, None
    try:
        d = json.load(open(p))
        for k, v in d.get("kernels", {}).items():
            if kernel_substr in k:
                return float(v["dram_bytes_per_launch"]), "profiles/r02_roofline_traffic.json (%s; ncu --set full, %s)" % (k, d.get("source", "?"))
    except Exception:
        pass
    return None, None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms (a timed region of a few
    tens of milliseconds still yields a median), nvidia-smi every 100 ms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
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
                    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                    mask = sum(self.BITS[n] for i, n in enumerate(names) if f[2 + i].lower().startswith("active"))
                    self.samples.append((int(f[0]), int(f[1]), mask))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self, lo=0, hi=None):
        smp = self.samples[lo:hi]
        sm = sorted(s[0] for s in smp)
        mx = [s[1] for s in smp]
        reasons = sorted({n for s in smp for n, bit in self.BITS.items() if s[2] & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(smp), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------------
# synthetic models / inputs
# ------------------------------------------------------------------------------------------------------------------

def _calibrate_3d(m):
    """BatchNorm running variance = the analytic output variance of its He-initialised conv, so that eval-mode
    activations stay O(1) (random-init weights of the reference architecture, stackhourglass.py:73-114)."""
    for mod in m.modules():
        if isinstance(mod, torch.nn.Sequential) and len(mod) >= 2 and isinstance(mod[1], torch.nn.BatchNorm3d):
            w = mod[0].weight
            tr = isinstance(mod[0], torch.nn.ConvTranspose3d)
            cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
            if tr:
                mod[0].weight.data.normal_(0, (2.0 / (27 * cout)) ** 0.5)
                mod[1].running_var.fill_(cin / (4.0 * cout))      # 27/8 taps reach an output voxel on average
            else:
                mod[1].running_var.fill_(2.0 * cin / cout)


def synthetic_hotpath(device):
    from dsmnet_b200.psmnet import PSMNetHotPath
    torch.manual_seed(0)
    m = PSMNetHotPath(MAXDISP)
    _calibrate_3d(m)
    return m.to(device).eval()


def synthetic_psmnet(device):
    from dsmnet_b200.psmnet import PSMNet
    torch.manual_seed(0)
    m = PSMNet(MAXDISP)
    _calibrate_3d(m)
    for mod in m.feature_extraction.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_var.fill_(2.0)
    return m.to(device).eval()


# ------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port) — the only place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------------------------

def cpu_sample(workload, H, W, reps=1, threads=None):
    """(pairs/s, seconds per pass, threads): the oracle port of the workload's path on one full pair on the host cores."""
    import oracle.ops as O
    torch.set_num_threads(threads or os.cpu_count())
    g = torch.Generator().manual_seed(0)
    h, w = H // 4, W // 4
    best = None
    with torch.no_grad():
        if workload == "psmnet_full_384x1248":
            from dsmnet_b200.psmnet import feature_extraction
            shapes = {k: tuple(v.shape) for k, v in feature_extraction().state_dict().items()}
            tp = O.trunk_random_params(shapes, seed=0)
            imL = torch.rand(1, 3, H, W, generator=g); imR = torch.rand(1, 3, H, W, generator=g)
        else:
            fL = torch.randn(1, C_FEAT, h, w, generator=g); fR = torch.randn(1, C_FEAT, h, w, generator=g)
        params = O.psmnet_random_params(seed=0)      # timing does not depend on the values
        for _ in range(reps):
            t0 = time.perf_counter()
            if workload == "psmnet_full_384x1248":
                fL = O.psmnet_feature_extraction(tp, imL); fR = O.psmnet_feature_extraction(tp, imR)
            O.psmnet_hotpath(params, fL, fR, MAXDISP, (H, W))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return 1.0 / best, best, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    wl = args.workload
    if not wl.startswith("psmnet"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm is implemented for the PSMNet workloads (the metric's configuration); "
                          "workload %s has per-op CPU figures in SURVEY.md section 6" % wl}), flush=True)
        return
    H, W = (540, 960) if wl == "psmnet_540x960" else (384, 1248)
    times = []
    nthreads = os.cpu_count()
    for i in range(args.warmup + args.steps):
        v, dt, nthreads = cpu_sample(wl, H, W)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = 1.0 / (ms / 1e3)
    sample = "one full %dx%d pair per step through the oracle port of the path, fp32, torch CPU ops, %d threads" % (H, W, nthreads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl), "batch_per_gpu": args.batch, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(wl):
    return {"psmnet_384x1248": "psmnet_hotpath_384x1248_maxdisp192", "psmnet_540x960": "psmnet_hotpath_540x960_maxdisp192",
            "psmnet_full_384x1248": "psmnet_whole_model_384x1248_maxdisp192"}.get(wl, wl)


# ------------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------------

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


def capture(step, no_graph=False):
    """(run, graph, static_out): the step captured once in a CUDA graph (eager launches are the same kernels)"""
    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    graph = None
    if not no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):    # NCCL's watchdog thread may touch CUDA meanwhile
                out = step()
        except Exception as e:
            graph = None
            sys.stderr.write("CUDA graph capture unavailable (%s); timing eager launches\n" % e)
    return ((lambda: graph.replay()) if graph is not None else step), graph, out


def stock_torch_hotpath(m, fL, fR, maxdisp, out_hw):
    """The SAME path as the reference writes it, on stock PyTorch CUDA ops (cuDNN conv3d, F.interpolate, softmax) with the
    module's own nn layers — stackhourglass.py:123-168.  Context for the hand-written kernels, not a product path."""
    import torch.nn.functional as F
    B, C, h, w = fL.shape
    D = maxdisp // 4
    cost = fL.new_zeros(B, 2 * C, D, h, w)
    for i in range(D):
        if i > 0:
            cost[:, :C, i, :, i:] = fL[:, :, :, i:]
            cost[:, C:, i, :, i:] = fR[:, :, :, :-i]
        else:
            cost[:, :C, i] = fL; cost[:, C:, i] = fR
    if fL.dtype == torch.bfloat16:
        cost = cost.contiguous(memory_format=torch.channels_last_3d)

    def add(a, b):
        d, hh, ww = (min(x, y) for x, y in zip(a.shape[2:], b.shape[2:]))
        return a[:, :, :d, :hh, :ww] + b[:, :, :d, :hh, :ww]

    def hg(hm, x, presqu, postsqu):
        out = hm.conv1(x)
        pre = hm.conv2(out)
        pre = F.relu(add(pre, postsqu) if postsqu is not None else pre)
        out = hm.conv4(hm.conv3(pre))
        post = F.relu(add(hm.conv5(out), presqu if presqu is not None else pre))
        return hm.conv6(post), pre, post

    cost0 = m.dres0(cost)
    cost0 = add(m.dres1(cost0), cost0)
    out1, pre1, post1 = hg(m.dres2, cost0, None, None); out1 = add(out1, cost0)
    out2, pre2, post2 = hg(m.dres3, out1, pre1, post1); out2 = add(out2, cost0)
    out3, pre3, post3 = hg(m.dres4, out2, pre1, post2); out3 = add(out3, cost0)
    c1 = m.classif1(out1); c2 = m.classif2(out2) + c1; c3 = m.classif3(out3) + c2
    preds = []
    disp = torch.arange(maxdisp, device=fL.device, dtype=torch.float32)
    for c in (c3, c2, c1):
        up = F.interpolate(c.float(), [maxdisp, out_hw[0], out_hw[1]], mode="trilinear", align_corners=True).squeeze(1)
        preds.append(F.softmax(up, dim=1).permute(0, 2, 3, 1).matmul(disp))
    return preds


def torch_gpu_baseline(m, fL, fR, H, W, steps=3):
    """pairs/s of stock_torch_hotpath on this GPU: fp32 with TF32 off (the reference's arithmetic) and bf16 autocast with
    channels_last_3d weights/activations (the fastest stock configuration cuDNN offers for these layers)."""
    res = {}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    B = fL.shape[0]
    try:
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            ms = time_kernel_alone(lambda: stock_torch_hotpath(m, fL, fR, MAXDISP, (H, W)), reps=steps)
        res["fp32_no_tf32"] = {"pairs_per_s": B / (ms / 1e3), "ms_per_step": ms}
        torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
        from dsmnet_b200.psmnet import PSMNetHotPath
        m2 = PSMNetHotPath(MAXDISP)
        m2.load_state_dict(m.state_dict())
        m2 = m2.to(fL.device).eval().to(memory_format=torch.channels_last_3d)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ms = time_kernel_alone(lambda: stock_torch_hotpath(m2, fL.bfloat16(), fR.bfloat16(), MAXDISP, (H, W)), reps=steps)
        res["bf16_autocast_channels_last_3d"] = {"pairs_per_s": B / (ms / 1e3), "ms_per_step": ms}
        del m2
    except Exception as e:
        res["error"] = str(e)[:200]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.cuda.empty_cache()
    res["what"] = "the same path (volume loop, nn.Conv3d/BatchNorm3d/ReLU modules, F.interpolate + softmax heads) on stock PyTorch %s / cuDNN, eager, this GPU" % torch.__version__
    return res


# ------------------------------------------------------------------------------------------------------------------
# the PSMNet workloads (hot path / whole model)
# ------------------------------------------------------------------------------------------------------------------

def run_psmnet(args, rank, world, device, dist, barrier):
    from dsmnet_b200.conv3d import conv_timeouts
    wl = args.workload
    full = wl == "psmnet_full_384x1248"
    H, W = (540, 960) if wl == "psmnet_540x960" else (384, 1248)
    B = args.batch
    h, w = (H + 3) // 4, (W + 3) // 4
    m = synthetic_psmnet(device) if full else synthetic_hotpath(device)
    g = torch.Generator().manual_seed(1000 + rank)
    if full:
        host_in = [torch.rand(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    else:
        host_in = [torch.randn(B, C_FEAT, h, w, generator=g).pin_memory() for _ in range(2)]
    dev_in = [t.to(device) for t in host_in]
    trunk_launches = 59 if full else 0
    launches_per_step = trunk_launches + 1 + 28 + 3      # (trunk,) concat volume, 28 conv blocks, three head launches

    def step():
        with torch.no_grad():
            return m(dev_in[0], dev_in[1], "test")[1] if full else m(dev_in[0], dev_in[1], (H, W))

    run, graph, static_out = capture(step, args.no_graph)
    for _ in range(max(args.warmup, 3)):
        run()
    sampler = ClockSampler(device.index or 0); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sampler.active = True
    if args.profile_range:                       # ncu --profile-from-start off: only the timed steps are captured
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    sampler.active = False
    t_ms = e0.elapsed_time(e1)
    n_burst = len(sampler.samples)

    # ---- sustained: >= 2 s of back-to-back steps (the SM clock settles under power; the burst above does not see that) ----
    sustained = None
    if not args.no_sustained:
        est = max(t_ms / args.steps, 1e-3)
        n_sus = int(max(args.sustained_seconds * 1e3 / est, args.steps))
        barrier()
        sampler.active = True
        e0.record()
        for _ in range(n_sus):
            run()
        e1.record()
        barrier()
        sampler.active = False
        t_sus = e0.elapsed_time(e1)
        sustained = {"steps": n_sus, "seconds": t_sus / 1e3, "ms_per_step": t_sus / n_sus, "clocks": sampler.summary(n_burst)}
    n_sus_samples = len(sampler.samples)

    # ---- end to end through the package's host-buffer API -----------------------------------------------------------
    pipe = m.image_pipeline(B, H, W, device, not args.no_graph) if full else m.host_pipeline(B, h, w, (H, W), device, not args.no_graph)

    def e2e_run(n):
        for i in range(n):
            t = pipe.submit(*host_in)
            if i >= 1:
                pipe.wait(t - 1)                                   # the caller reads the result of step i-1
        pipe.wait(pipe.count - 1)

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
        vals = [t_ms, t_e2e * 1e3, sustained["ms_per_step"] * sustained["steps"] if sustained else 0.0]
        tt = torch.tensor(vals, device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, t_e2e = float(tt[0]), float(tt[1]) / 1e3
        if sustained:
            sustained["seconds"] = float(tt[2]) / 1e3; sustained["ms_per_step"] = float(tt[2]) / sustained["steps"]
    timeouts = conv_timeouts()
    if rank != 0:
        return

    hbm, tf_burst, tf_sust, which = measured_peaks()
    if sustained:
        sustained["value"] = B * world * sustained["steps"] / sustained["seconds"]
        sustained["unit"] = "pairs/s"
        if (H, W) == (384, 1248):
            tfs = B * 926.7e9 / (sustained["ms_per_step"] * 1e-3) / 1e12
            sustained["stack_tflops"] = tfs
            sustained["stack_frac_of_sustained_peak"] = tfs / tf_sust
            sustained["note"] = "926.7 GFLOP of 3-D convolutions per pair over the whole step (volume, heads%s included in the time) vs the measured sustained bf16 peak" % (", 2-D trunk" if full else "")

    # ---- roofline of the dominant kernel (6 of the 28 conv launches: Conv3d 32->32 @ D/4 x H/4 x W/4) -----------------
    hot = m
    plan = hot._get_plan(device)
    ws = hot._workspace(B, MAXDISP // 4, h, w, device)
    layer = plan.dres0_2
    ms_k = time_kernel_alone(lambda: layer(ws["a"], ws["c0"]))
    flops = 2.0 * 27 * 32 * 32 * B * (MAXDISP // 4) * h * w
    achieved = flops / (ms_k * 1e-3) / 1e12
    traffic, tsrc = profiled_traffic("conv3d_rs_kernel<32, 32, 0>")
    if traffic is not None and (B, H, W) != (1, 384, 1248):
        traffic, tsrc = None, "captured at batch 1, 384x1248 only"
    roofline = {"bound": "tensor", "kernel": "conv3d_rs_kernel<KC=32,NP=32> (Conv3d 32->32 k3 s1 + BN + ReLU, 6 launches/step)",
                "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                "peak_source": which + " bf16 burst (kernel timed alone)", "ms_per_launch": ms_k,
                "frac_of_sustained_peak": achieved / tf_sust, "traffic": traffic, "traffic_source": tsrc}

    cpu = None
    gpu_stock = None
    whole = None
    if world == 1 and not args.no_cpu_baseline:
        v, dt, nthreads = cpu_sample(wl, H, W, reps=2 if full else 3)
        cpu = {"value": v, "unit": "pairs/s", "cores": nthreads, "kind": "port",
               "sample": "oracle port (fp32 torch CPU ops) of the whole path on one full %dx%d pair, best of %d passes of %.1f s" %
                         (H, W, 2 if full else 3, dt)}
    if world == 1 and not full and not args.no_torch_baseline and B <= 2:
        gpu_stock = torch_gpu_baseline(m, dev_in[0], dev_in[1], H, W)
    if world == 1 and wl == "psmnet_384x1248" and not args.no_whole_model and B == 1:
        try:
            del pipe
            wm = synthetic_psmnet(device)
            imgs = [torch.rand(1, 3, H, W, device=device) for _ in range(2)]
            with torch.no_grad():
                wrun, wgraph, _ = capture(lambda: wm(imgs[0], imgs[1], "test")[1], args.no_graph)
            ms_w = time_kernel_alone(wrun, reps=20)
            with torch.no_grad():
                trun, _, _ = capture(lambda: wm.feature_extraction(torch.cat(imgs, 0)), args.no_graph)
            ms_t = time_kernel_alone(trun, reps=20)
            whole = {"what": "whole drop-in PSMNet from images (2-D trunk on the library's kernels, both images as one batch, + the hot path), one CUDA graph",
                     "pairs_per_s": 1e3 / ms_w, "ms_per_pair": ms_w, "trunk_ms": ms_t, "trunk_tflops": 424.0e9 / (ms_t * 1e-3) / 1e12}
        except Exception as e:
            whole = {"error": str(e)[:300]}

    pairs = B * world * args.steps

    def nbytes(v):
        t = v if isinstance(v, torch.Tensor) else v.data
        return t.numel() * t.element_size()
    act_bytes = sum(nbytes(v) for k, v in ws.items() if v is not None and not isinstance(v, list)) + \
        sum(nbytes(x) for k in ("out", "pre", "post") for x in ws[k])
    line = {"metric": METRIC, "value": pairs / (t_ms / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(wl), "batch_per_gpu": B,
                       "volume": "64x%dx%dx%d per pair" % (MAXDISP // 4, h, w), "sharding": "stereo pairs over ranks, no data-path collective",
                       "l2": "per-step activation working set %.2f GB >> 126 MB L2 (no explicit flush)" % (act_bytes / 1e9),
                       "cuda_graph": graph is not None, "conv_timeouts": timeouts,
                       "inputs": "images [B,3,%d,%d] fp32" % (H, W) if full else "feature maps [B,32,%d,%d] fp32" % (h, w)},
            "clocks": sampler.summary(0, n_burst),
            "sustained": sustained,
            "e2e": {"value": pairs / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": sum(t.numel() * 4 for t in host_in),
                    "d2h_bytes_per_step": 3 * B * H * W * 4,
                    "api": "PSMNet.image_pipeline" if full else "PSMNetHotPath.host_pipeline", "clocks": sampler.summary(n_sus_samples)},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "torch_gpu_baseline": gpu_stock, "whole_model": whole}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs 3, 4, 5b
# ------------------------------------------------------------------------------------------------------------------

def _timed_steps(run, args, device, dist, barrier):
    for _ in range(max(args.warmup, 3)):
        run()
    sampler = ClockSampler(device.index or 0); sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sampler.active = True
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    sampler.active = False
    sampler.stop_flag = True; sampler.join(timeout=2)
    t_ms = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([t_ms], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt[0])
    return t_ms, sampler.summary()


def run_gcnet(args, rank, world, device, dist, barrier):
    """GC-Net 3-D path, forward + backward with train-mode BatchNorm on one 256x512 crop per GPU (BASELINE configs[2])."""
    from dsmnet_b200.gcnet import GCNetHotPath
    from dsmnet_b200.conv3d import conv_timeouts
    torch.manual_seed(0)
    B = args.batch
    m = GCNetHotPath(MAXDISP).to(device).train()
    g = torch.Generator().manual_seed(1000 + rank)
    fL = torch.randn(B, 32, 128, 256, generator=g).to(device).requires_grad_()
    fR = torch.randn(B, 32, 128, 256, generator=g).to(device).requires_grad_()
    gt = (torch.rand(B, 1, 256, 512, generator=g) * 96).to(device)

    def step():
        m.zero_grad(set_to_none=True); fL.grad = fR.grad = None
        (m(fL, fR) - gt).abs().mean().backward()

    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    run, graph, _ = capture(step, args.no_graph)
    t_ms, clocks = _timed_steps(run, args, device, dist, barrier)
    if rank != 0:
        return
    _, tf_burst, tf_sust, which = measured_peaks()
    flops = 3 * 882.6e9 * B
    ms = t_ms / args.steps
    line = {"metric": "GC-Net 256x512 crops/sec, 3-D path forward+backward (train-mode BatchNorm)", "value": B * world * args.steps / (t_ms / 1e3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "gcnet_256x512_maxdisp192_fwd_bwd", "batch_per_gpu": B, "volume": "64x96x128x256 per crop",
                       "cuda_graph": graph is not None, "conv_timeouts": conv_timeouts(), "l2": "activations >> 126 MB L2"},
            "clocks": clocks, "e2e": None, "gpu_launches": None,
            "roofline": {"bound": "tensor", "kernel": "whole step (fwd + dgrad + wgrad convolutions, 2.65 TFLOP algorithmic, BatchNorm streams included in the time)",
                         "achieved": flops / (ms * 1e-3) / 1e12, "peak": tf_sust, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / tf_sust,
                         "peak_source": which + " bf16 sustained", "traffic": None},
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def run_iresnet_ops(args, rank, world, device, dist, barrier):
    """iResNet's hot-path ops at the SceneFlow shape (BASELINE configs[3]): Corr1d D=81 on 128x135x240, the imwrap feature
    warp of 32x540x960, Corr1d(k3, s2) D=41 on 64x270x480 — one launch chain per pair."""
    from dsmnet_b200.corr1d import corr1d
    from dsmnet_b200.imwrap import imwrap_BCHW
    import torch.nn.functional as F
    B = args.batch
    g = torch.Generator().manual_seed(1000 + rank)
    a1 = torch.relu(torch.randn(B, 128, 135, 240, generator=g)).to(device); b1 = torch.relu(torch.randn(B, 128, 135, 240, generator=g)).to(device)
    a2 = torch.relu(torch.randn(B, 64, 270, 480, generator=g)).to(device); b2 = torch.relu(torch.randn(B, 64, 270, 480, generator=g)).to(device)
    src = torch.rand(B, 32, 540, 960, generator=g).to(device); disp = (torch.rand(B, 1, 540, 960, generator=g) * 96).to(device)

    def step():
        with torch.no_grad():
            c1 = corr1d(a1, b1, 81, 1)
            wv = imwrap_BCHW(src, disp, delt=5e-5)
            c2 = F.avg_pool2d(corr1d(a2, b2, 41, 2), 3, stride=1, padding=1)
            return c1, wv, c2

    run, graph, _ = capture(step, args.no_graph)
    t_ms, clocks = _timed_steps(run, args, device, dist, barrier)
    if rank != 0:
        return
    hbm, _, _, which = measured_peaks()
    bytes_step = B * 4.0 * (135 * 240 * (2 * 128 + 81) + 270 * 480 * (2 * 64 + 41) + (32 * 540 * 960 * 2 + 540 * 960))
    ms = t_ms / args.steps
    line = {"metric": "iResNet hot-path ops (Corr1d x2 + imwrap) pairs/sec at 540x960", "value": B * world * args.steps / (t_ms / 1e3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "iresnet_ops_540x960", "batch_per_gpu": B, "cuda_graph": graph is not None,
                       "l2": "inputs %.0f MB per pair > 126 MB L2" % (bytes_step / B / 1e6)},
            "clocks": clocks, "e2e": None, "gpu_launches": 4 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "the op chain (algorithmic bytes of the three ops / step time)", "achieved": bytes_step / (ms * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": bytes_step / (ms * 1e-3) / 1e9 / hbm, "peak_source": which + " HBM copy", "traffic": None},
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def run_selfsup(args, rank, world, device, dist, barrier):
    """Self-supervised DispNetC training step (stereo_selfsupervised.py:60-100, `depthmono-mask`; BASELINE configs[4b]):
    two model forwards, 28 imwrap warps, SSIM / smoothness terms, Adam; DistributedDataParallel (32 MB buckets) carries the
    168.7 MB gradient all-reduce over NCCL when N > 1."""
    from dsmnet_b200.dispnetcorr import dispnetcorr
    from dsmnet_b200.selfsup import train_step
    torch.manual_seed(0)
    B = args.batch
    model = dispnetcorr(192).to(device).train()
    nparams = sum(p.numel() for p in model.parameters())
    if dist is not None:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], bucket_cap_mb=32, gradient_as_bucket_view=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1000 + rank)
    batch = torch.rand(B, 6, 384, 768, generator=g).to(device)
    wlv = [1.0, 0.01, 0.01, 0.01, 0.01, 0.01, 0.01]
    t_ms, clocks = _timed_steps(lambda: train_step(model, opt, batch, 64, wlv), args, device, dist, barrier)
    if rank != 0:
        return
    ms = t_ms / args.steps
    line = {"metric": "self-supervised DispNetC training pairs/sec (768x384 crops, nedge 64)", "value": B * world * args.steps / (t_ms / 1e3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "dispnetc_selfsup_train_step_768x384_nedge64", "batch_per_gpu": B, "params": nparams,
                       "collective": "NCCL all-reduce of %.1f MB fp32 gradients per step (DDP, 32 MB buckets, overlapped with backward)" % (nparams * 4 / 1e6)
                       if dist is not None else "none (1 GPU)", "l2": "activations >> 126 MB L2"},
            "clocks": clocks, "e2e": None, "gpu_launches": None, "roofline": None, "cpu_baseline": None}
    print(json.dumps(line), flush=True)


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
    _lib.lib()                                  # fail loudly if the CUDA library is missing

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    wl = args.workload
    try:
        if wl.startswith("psmnet"):
            run_psmnet(args, rank, world, device, dist, barrier)
        elif wl == "gcnet_256x512":
            run_gcnet(args, rank, world, device, dist, barrier)
        elif wl == "iresnet_ops_540x960":
            run_iresnet_ops(args, rank, world, device, dist, barrier)
        else:
            run_selfsup(args, rank, world, device, dist, barrier)
    finally:
        if dist is not None:
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="psmnet_384x1248", choices=WORKLOADS)
    ap.add_argument("--batch", type=int, default=None, help="stereo pairs per GPU per step (default 1; 4 for the training workload)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-whole-model", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.batch is None:      # the op chain of iResNet takes 0.2 ms per pair: 8 pairs per step give the clock sampler a region to see
        args.batch = {"dispnetc_selfsup_train": 4, "iresnet_ops_540x960": 8}.get(args.workload, 1)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
