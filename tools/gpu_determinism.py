"""Bitwise repeatability of the team kernels (a shared-memory race would show as run-to-run differences)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
for name, B in (("cfg3_thing_box_arch", 1024), ("cfg5_thing_robust8", 1024)):
    desc, meta = workload.load(name)
    mpc = BatchedMPC(desc, "f32")
    ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
    b = workload.sample_batch(name, desc, meta, B, 7, ee)
    outs = [mpc.solve(b["x0"], b["target"], b["body_params"]) for _ in range(4)]
    same = all(np.array_equal(outs[0]["X"], o["X"], equal_nan=True) and np.array_equal(outs[0]["U"], o["U"], equal_nan=True)
               and np.array_equal(outs[0]["stats"], o["stats"], equal_nan=True) for o in outs[1:])
    print(name, "bitwise identical over 4 runs:", same, "status", [int((outs[0]["status"] == s).sum()) for s in range(4)])
