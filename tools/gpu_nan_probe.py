"""Instances whose fp32 GPU solve ends with status NAN: what do the fp64 GPU kernel and the oracle say?"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
for name, B in [("cfg4_thing_obstacles2", 2048), ("cfg3_thing_box_arch", 2048), ("cfg1_ur10_demo", 4096)]:
    desc, meta = workload.load(name)
    m32, m64 = BatchedMPC(desc, "f32"), BatchedMPC(desc, "f64")
    ee = lambda x: m64.eval("end_effector_position", x, np.zeros((x.shape[0], m64.nu)))
    mg = (lambda x: m64.eval("obstacle_avoidance", x, np.zeros((x.shape[0], m64.nu)))) if desc.obstacles_enabled else None
    b = workload.sample_batch(name, desc, meta, B, 1234, ee, margin_fn=mg)
    o32 = m32.solve(b["x0"], b["target"], b["body_params"])
    bad = np.nonzero(o32["status"] != 0)[0]
    print(name, "fp32 non-converged:", len(bad), "status", o32["status"][bad][:20])
    if len(bad) == 0:
        continue
    idx = bad[:16]
    bp = None if b["body_params"] is None else b["body_params"][idx]
    o64 = m64.solve(b["x0"][idx], b["target"][idx], bp)
    orc = oracle.solve_batch(desc, b["x0"][idx], b["target"][idx], bp)
    print("  fp64 GPU status", o64["status"], "iters", o64["stats"][:, 0])
    print("  oracle   status", orc["status"], "iters", orc["stats"][:, 0])
    print("  fp32 iters", o32["stats"][idx, 0], "reason", o32["stats"][idx, 3], "last step", o32["stats"][idx, 4])
    if mg is not None:
        print("  start margin", mg(b["x0"][idx]).min(axis=1))
