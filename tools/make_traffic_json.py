#!/usr/bin/env python3
"""profiles/traffic.json (read by bench.py: roofline.traffic) from the raw metric pages of the committed
`ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum of one launch of the solve kernel.
usage: python tools/make_traffic_json.py <tag>   (reads profiles/<tag>_<cfgN>_raw.csv, e.g. tag = r2_v9)"""
import csv
import json
import sys
from pathlib import Path

root = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
names = {"cfg1": ("cfg1_ur10_demo", 4096), "cfg2": ("cfg2_thing_demo", 4096), "cfg3": ("cfg3_thing_box_arch", 4096),
         "cfg4": ("cfg4_thing_obstacles2", 2048), "cfg5": ("cfg5_thing_robust8", 1024)}
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
path = root / "profiles" / "traffic.json"
out = json.load(open(path)) if path.exists() else {}   # entries of other captures are kept
for short, (name, batch) in names.items():
    f = root / "profiles" / f"{tag}_{short}_raw.csv"
    if not f.exists():
        continue
    rows = list(csv.reader(open(f)))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    rd = float(d["dram__bytes_read.sum"][0]) * unit[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0]) * unit[d["dram__bytes_write.sum"][1]]
    ms = float(d["gpu__time_duration.sum"][0]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[d["gpu__time_duration.sum"][1]]
    out[name] = {"batch": batch, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "kernel_ms_under_ncu": ms,
                 "source": f"profiles/{f.name} (ncu --set full, one launch of ub::solve_batch_kernel, {short}, B={batch})"}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps({k: round(v["dram_bytes_per_launch"] / 1e9, 3) for k, v in out.items()}))
