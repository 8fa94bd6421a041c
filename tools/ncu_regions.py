#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per source function and per line:
stall samples by reason and executed warp instructions.
usage: python tools/ncu_regions.py gpurun_out/prof.ncu-rep [--lines N]"""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
root = Path(__file__).resolve().parent.parent
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))


def functions(path):
    out = []
    pat = re.compile(r"__device__[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, l in enumerate(path.read_text().splitlines(), 1):
        m = pat.search(l)
        if m and not l.strip().startswith("//"):
            out.append((i, m.group(1)))
    return out


tables = {f: functions(root / "upright_b200" / "csrc" / f) for f in ("ub_solver.cuh", "ub_device.cuh")}


def region(short, line):
    if short not in tables:
        return short
    name = "(top)"
    for start, fn in tables[short]:
        if start <= line:
            name = fn
        else:
            break
    return f"{short}:{name}"


STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_selected", "stall_not_selected",
          "stall_branch_resolving", "stall_barrier", "stall_mio", "stall_lg", "stall_math", "stall_dispatch"]
cur, hdr = None, None
reg = collections.defaultdict(lambda: collections.Counter())
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if r[0] == "Function Name" or hdr is None or not r[0].isdigit():
        continue

    def g(k):
        try:
            return float(r[hdr[k]])
        except Exception:
            return 0.0
    key = region(cur, int(r[0]))
    c = reg[key]
    c["samples"] += g("# Samples")
    c["inst"] += g("Instructions Executed")
    for s in STALLS:
        c[s] += g(s)
    lines.append((g("# Samples"), cur, int(r[0]), r[1].strip()[:80], g("Instructions Executed"), g("stall_long_sb"), g("stall_no_inst")))
tot = sum(c["samples"] for c in reg.values())
toti = sum(c["inst"] for c in reg.values())
print(f"total samples {tot:.0f}, warp instructions {toti / 1e9:.3f} G")
print(f"{'region':45s} {'smp%':>6s} {'inst%':>6s} | " + " ".join(f"{s[6:10]:>5s}" for s in STALLS))
for k, c in sorted(reg.items(), key=lambda kv: -kv[1]["samples"])[:32]:
    print(f"{k:45s} {100 * c['samples'] / tot:6.2f} {100 * c['inst'] / toti:6.2f} | " +
          " ".join(f"{100 * c[s] / tot:5.2f}" for s in STALLS))
print("all", " " * 41, f"{100:6.2f} {100:6.2f} | " + " ".join(f"{100 * sum(c[s] for c in reg.values()) / tot:5.2f}" for s in STALLS))
print()
lines.sort(reverse=True)
for a in lines[:nlines]:
    print(f"{100 * a[0] / tot:5.2f}% inst {100 * a[4] / toti:5.2f}% long_sb {a[5]:7.0f} no_inst {a[6]:6.0f} | {a[1]}:{a[2]} {a[3]}")
