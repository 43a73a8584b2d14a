#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an ncu report -> profiles/r02_roofline_traffic.json, which bench.py reads for
`roofline.traffic` (run here, no GPU needed):

    python tools/ncu_traffic.py gpurun_out/X.ncu-rep "description of the capture" [out.json]

traffic = dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, averaged over the captured
launches of the same kernel (template arguments kept, parameter list dropped)."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, source = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    acc = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*$", "", d.get("Kernel Name", "?")).strip()
        f = lambda k: float((d.get(k) or "0").replace(",", "")) * UNIT.get(u.get(k, "byte"), 1.0)
        e = acc.setdefault(name, {"launches": 0, "bytes": 0.0, "ns": 0.0})
        e["launches"] += 1
        e["bytes"] += f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
        dur = float((d.get("gpu__time_duration.sum") or "0").replace(",", ""))
        e["ns"] += dur * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u.get("gpu__time_duration.sum", "ns"), 1.0)
    res = {"source": source, "report": os.path.basename(rep),
           "kernels": {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"],
                           "ncu_us_per_launch": v["ns"] / v["launches"] / 1e3} for k, v in acc.items()}}
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
