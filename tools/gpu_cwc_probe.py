"""Contact-wrench-cone violation of the planned knots of cfg2 over successive SQP iterations (upright_b200/robust.py)."""
import sys, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from upright_b200 import workload, problem_io
from upright_b200.engine import BatchedMPC
from upright_b200.robust import WrenchConeVerifier
name="cfg2_thing_demo"
desc, meta = problem_io.load_fixture(name)
mpc = BatchedMPC(desc, "f32")
B=64
b = workload.sample_batch(name, desc, meta, B, 5, lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu))), vary_bodies=False)
ver = WrenchConeVerifier(mpc, desc)
out = mpc.solve(b["x0"], b["target"], None)
for it in range(4):
    V = ver.violation(out["X"][:, :-1])
    ok = out["status"]==0
    print(it, "ok", ok.sum(), "median of per-instance max %.3e  p90 %.3e worst %.3e; knot0 max %.3e; knots>=1 median %.3e" % (np.median(V[ok].max(1)), np.percentile(V[ok].max(1),90), V[ok].max(), V[ok][:,0].max(), np.median(V[ok][:,1:].max(1))), "viol stat", np.median(out["stats"][:,2]))
    out = mpc.solve(b["x0"], b["target"], None, X=out["X"], U=out["U"], warm=True)
