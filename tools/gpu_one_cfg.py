"""A few launches of the solve kernel of ONE configuration (for ncu captures of the other kernel variants)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
name, B = sys.argv[1], int(sys.argv[2])
desc, meta = workload.load(name)
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
mg = (lambda x: mpc.eval("obstacle_avoidance", x, np.zeros((x.shape[0], mpc.nu)))) if desc.obstacles_enabled else None
b = workload.sample_batch(name, desc, meta, B, 1234, ee, margin_fn=mg)
dev = lambda a: None if a is None else torch.tensor(a, dtype=torch.float32, device="cuda")
x0, tg, bp = dev(b["x0"]), dev(b["target"]), dev(b["body_params"])
for _ in range(3):
    out = mpc.solve_device(x0, tg, bp)
    torch.cuda.synchronize()
    print(name, B, "ms", mpc.last_solve_ms())
