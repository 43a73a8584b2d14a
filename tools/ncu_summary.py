#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): key raw metrics per captured launch and the
top stall sites of the first launch.   python tools/ncu_summary.py report.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:                                  # count the launches of each (kernel, grid); print the first of each
        d = dict(zip(hdr, r))
        key = (d.get("Kernel Name"), d.get("launch__grid_size"))
        seen[key] = seen.get(key, 0) + 1
    done = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        key = (d.get("Kernel Name"), d.get("launch__grid_size"))
        if key in done:
            continue
        done.add(key)
        print("== %s   [%d captured launch(es) with this grid; first shown]" % (d.get("Kernel Name"), seen[key]))
        for k in KEYS:
            if k in d:
                print("   %-95s %s %s" % (k, d[k], u.get(k, "")))
        rd = float(d.get("dram__bytes_read.sum", "0").replace(",", "") or 0); wr = float(d.get("dram__bytes_write.sum", "0").replace(",", "") or 0)
        print("   traffic = dram read + write = %.1f %s" % (rd + wr, u.get("dram__bytes_read.sum", "")))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) < 3:
        return
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:                                   # first launch only: the page repeats per launch
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr) and r[ix["# Samples"]].isdigit():
            data.append(r)
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print("-- warp-state samples of the first launch: %d; top stall sites (SASS, samples, dominant reason)" % tot)
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
        st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stall), key=lambda x: -x[1])[:1]
        print("   %5s  %-70s %s" % (r[ix["# Samples"]], r[ix["Source"]][:70], st))


if __name__ == "__main__":
    main()
