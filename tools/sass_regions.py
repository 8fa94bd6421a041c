#!/usr/bin/env python3
"""SASS instruction count per source function (by line range of the enclosing `__device__` function) of one
kernel object — the kernel is instruction-fetch sensitive, so code size per phase is a tracked quantity.
usage: python tools/sass_regions.py build/obj/ub_kernel_thing_1obj_f32.o"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

obj = Path(sys.argv[1]).resolve()
root = Path(__file__).resolve().parent.parent
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", str(obj)], cwd=td, stdout=subprocess.DEVNULL)
    cubin = next(Path(td).glob("*.cubin"))
    sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout

# function line ranges of the two kernel headers
def functions(path):
    out, lines = [], path.read_text().splitlines()
    pat = re.compile(r"__device__[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, l in enumerate(lines, 1):
        m = pat.search(l)
        if m and not l.strip().startswith("//"):
            out.append((i, m.group(1)))
    return out

tables = {}
for f in ("ub_solver.cuh", "ub_device.cuh"):
    tables[f] = functions(root / "upright_b200" / "csrc" / f)

def region(fname, line):
    short = fname.split("/")[-1]
    if short not in tables:
        return short
    name = "(top)"
    for start, fn in tables[short]:
        if start <= line:
            name = fn
        else:
            break
    return f"{short}:{name}"

cnt = collections.Counter()
cur = ("?", 0)
total = 0
for l in sass.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        cnt[region(*cur)] += 1
        total += 1
print(f"total SASS instructions: {total} ({total * 16 / 1024:.0f} KB)")
for k, v in cnt.most_common(40):
    print(f"{v:7d}  {k}")
