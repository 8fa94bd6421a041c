#!/usr/bin/env python3
"""Build a variant of the CUDA library with extra nvcc flags into variants/<name>/libupright_b200.so (select it
with UB_LIBRARY=...): A/B experiments on the GPU box without touching the in-tree library.
usage: python tools/build_variant.py <name> [-DFLAG=VALUE ...]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

root = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(root))
from __graft_entry__ import NVCC_FLAGS  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = root / "variants" / name
out_dir.mkdir(parents=True, exist_ok=True)
units = sorted((root / "upright_b200" / "csrc").glob("*.cu"))


def compile_unit(u):
    o = out_dir / (u.stem + ".o")
    subprocess.check_call(["nvcc", *NVCC_FLAGS, *extra, "-c", "-o", str(o), str(u)])
    return str(o)


with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
    objs = list(ex.map(compile_unit, units))
so = out_dir / "libupright_b200.so"
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(so), *objs])
for o in objs:
    os.remove(o)
print(so)
