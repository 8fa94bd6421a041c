"""Kernel time and status counts of every BASELINE configuration in both arithmetic widths (fp32 / fp64 kernels)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["f32", "f64"]
for name, B in [("cfg1_ur10_demo", 4096), ("cfg2_thing_demo", 4096), ("cfg3_thing_box_arch", 2048), ("cfg4_thing_obstacles2", 2048),
                ("cfg5_thing_robust8", 1024)]:
    desc, meta = workload.load(name)
    probe = BatchedMPC(desc, "f64")
    ee = lambda x: probe.eval("end_effector_position", x, np.zeros((x.shape[0], probe.nu)))
    mg = (lambda x: probe.eval("obstacle_avoidance", x, np.zeros((x.shape[0], probe.nu)))) if desc.obstacles_enabled else None
    b = workload.sample_batch(name, desc, meta, B, 1234, ee, margin_fn=mg)
    for prec in precs:
        mpc = BatchedMPC(desc, prec)
        dt = mpc.torch_dtype
        dev = lambda a: None if a is None else torch.tensor(a, dtype=dt, device="cuda")
        x0, tg, bp = dev(b["x0"]), dev(b["target"]), dev(b["body_params"])
        ts = []
        for _ in range(3):
            out = mpc.solve_device(x0, tg, bp)
            torch.cuda.synchronize()
            ts.append(mpc.last_solve_ms())
        it = out["stats"][:, 0].double()
        st = out["status"]
        print(f"{name:24s} {prec} B={B:5d} ms {min(ts[1:]):8.3f} solves/s {B / min(ts[1:]) * 1e3:9.0f} iters mean {it.mean():.2f} max {it.max():.0f} "
              f"status {[int((st == s).sum()) for s in range(4)]} smem/warp {mpc.layout()['s_total'] * (4 if prec == 'f32' else 8)} B "
              f"slot {mpc.layout()['total'] * (4 if prec == 'f32' else 8) // 1024} KB nan-reasons {sorted(set(out['stats'][:, 3][st == 3].cpu().numpy().astype(int).tolist()))}", flush=True)
