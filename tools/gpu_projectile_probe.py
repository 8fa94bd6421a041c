"""fp32 / fp64 kernels on the projectile-path problem of tests/_util.py: statuses and range-scaled errors against
the oracle for the warm-started, cold-started and flag-down solves (and the same problem with hard rows)."""
import copy
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle
from _util import ballistic_prediction, projectile_problem, projectile_throws
from upright_b200.engine import BatchedMPC

np.set_printoptions(linewidth=200, precision=3)
d, meta, tray = projectile_problem()
x0r = np.array(meta["x0"], dtype=float)
throws = projectile_throws(oracle.fk(d, np.concatenate((x0r, np.zeros(9))))["spheres"][tray])
Bn = len(throws)
x0 = np.hstack((np.tile(x0r, (Bn, 1)), throws))
target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (Bn, d.N + 1, 1))
Xw = np.stack([np.hstack((np.tile(x0r, (d.N + 1, 1)), ballistic_prediction(xo, d.N, d.dt))) for xo in throws])
Uw = np.zeros((Bn, d.N, 13))
rx = np.array(d.state_ub[:27]) - np.array(d.state_lb[:27])
ru = np.concatenate([np.array(d.input_ub[:9]) - np.array(d.input_lb[:9]), np.full(4, d.force_ub - d.force_lb)])


def run(desc, tag, warm):
    ref = oracle.solve_batch(desc, x0, target, X=Xw.copy() if warm else None, U=Uw.copy() if warm else None, warm=warm)
    for prec in ("f64", "f32"):
        m = BatchedMPC(desc, prec)
        o = m.solve(x0, target, None, X=Xw.copy() if warm else None, U=Uw.copy() if warm else None, warm=warm)
        ex = (np.abs(o["X"][:, :, :27] - ref["X"][:, :, :27]) / rx).reshape(Bn, -1).max(axis=1)
        eu = (np.abs(o["U"] - ref["U"]) / ru).reshape(Bn, -1).max(axis=1)
        print(f"{tag:28s} {prec} status {o['status']} ref {ref['status']} iters {o['stats'][:, 0]} ref {ref['stats'][:, 0]} ex {ex} eu {eu}", flush=True)


hard = copy.deepcopy(d)
hard.slacks.enabled = 0
down = copy.deepcopy(d)
down.projectile_active = 0.0
for desc, tag in ((d, "soft"), (down, "soft flag down"), (hard, "hard")):
    run(desc, tag + " warm", True)
    run(desc, tag + " cold", False)
