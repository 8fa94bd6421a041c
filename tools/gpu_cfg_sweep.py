"""Kernel time of every BASELINE configuration at (a fraction of) its batch size."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
import os
CASES = [("cfg1_ur10_demo", 4096), ("cfg2_thing_demo", 4096), ("cfg3_thing_box_arch", 4096), ("cfg4_thing_obstacles2", 2048),
         ("cfg5_thing_robust8", 1024)]
if os.environ.get("UB_SWEEP"):   # e.g. UB_SWEEP=cfg5_thing_robust8:8192,cfg3_thing_box_arch:4096
    CASES = [(c.split(":")[0], int(c.split(":")[1])) for c in os.environ["UB_SWEEP"].split(",")]
for name, B in CASES:
    desc, meta = workload.load(name)
    mpc = BatchedMPC(desc, "f32")
    ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
    mg = (lambda x: mpc.eval("obstacle_avoidance", x, np.zeros((x.shape[0], mpc.nu)))) if desc.obstacles_enabled else None
    b = workload.sample_batch(name, desc, meta, B, 1234, ee, margin_fn=mg)
    dev = lambda a: None if a is None else torch.tensor(a, dtype=torch.float32, device="cuda")
    x0, tg, bp = dev(b["x0"]), dev(b["target"]), dev(b["body_params"])
    ts = []
    for _ in range(4):
        out = mpc.solve_device(x0, tg, bp)
        torch.cuda.synchronize()
        ts.append(mpc.last_solve_ms())
    it = out["stats"][:, 0].double()
    st = out["status"]
    print(f"{name:24s} B={B:5d} ms {min(ts[1:]):8.3f} solves/s {B / min(ts[1:]) * 1e3:9.0f} iters mean {it.mean():.2f} max {it.max():.0f} "
          f"status {[int((st == s).sum()) for s in range(4)]} smem/warp {mpc.layout()['s_total'] * 4} B")
