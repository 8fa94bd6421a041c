"""Kernel time vs batch size (occupancy / wave effects)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_thing_demo"
desc, meta = workload.load(name)
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
full = workload.sample_batch(name, desc, meta, 8192, 1234, ee)
dev = lambda a: None if a is None else torch.tensor(a, dtype=torch.float32, device="cuda")
for B in [int(a) for a in sys.argv[2:]] or [296, 592, 1184, 2368, 4096, 4736, 8192]:
    x0, tg, bp = dev(full["x0"][:B]), dev(full["target"][:B]), dev(full["body_params"][:B])
    ts = []
    for _ in range(4):
        out = mpc.solve_device(x0, tg, bp)
        torch.cuda.synchronize()
        ts.append(mpc.last_solve_ms())
    it = out["stats"][:, 0].double()
    print(f"B={B:5d}  ms {min(ts[1:]):7.3f}  solves/s {B / min(ts[1:]) * 1e3:9.0f}  iters mean {it.mean():.2f} max {it.max():.0f}")
