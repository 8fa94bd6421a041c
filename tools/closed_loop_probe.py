"""Where the closed-loop time goes: feedback gains on/off, replans, plant steps."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
desc, meta = workload.load("cfg2_thing_demo")
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
b = workload.sample_batch("cfg2_thing_demo", desc, meta, 4096, 1234, ee)
goal = b["target"][:, :1, :]
for fb in (True, False):
    for n_steps, sim_dt, rp in ((100, 0.001, 0.01), (10, 0.01, 0.01), (100, 0.001, 0.05)):
        kw = dict(n_steps=n_steps, sim_dt=sim_dt, replan_period=rp, body_params=b["body_params"], log=False, use_feedback=fb)
        mpc.closed_loop(b["x0"], [0.0], goal, **kw)
        t0 = time.perf_counter()
        out = mpc.closed_loop(b["x0"], [0.0], goal, **kw)
        el = time.perf_counter() - t0
        print(f"feedback {fb!s:5} steps {n_steps:4d} dt {sim_dt} replan {rp}: {1e3 * el:7.2f} ms, {out['n_replans']} replans -> {1e3 * el / out['n_replans']:6.2f} ms/replan, last solve {mpc.last_solve_ms():.2f} ms")
