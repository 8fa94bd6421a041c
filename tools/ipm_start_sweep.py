"""CPU oracle: interior-point iterations of every BASELINE configuration against the start values of the slacks and
multipliers (qp_mu0, qp_thr0) — profiles/r2_v5_ipm_start_sweep.txt."""
import sys, numpy as np, itertools
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import oracle
from upright_b200 import workload
names = ["cfg1_ur10_demo","cfg2_thing_demo","cfg3_thing_box_arch","cfg4_thing_obstacles2","cfg5_thing_robust8"]
B = int(sys.argv[1]) if len(sys.argv)>1 else 32
for name in names:
    desc, meta = workload.load(name)
    ee = lambda x: oracle.fk_batch(desc, x)[:, :3] if hasattr(oracle,'fk_batch') else None
    try:
        batch = workload.sample_batch(name, desc, meta, B, 7, lambda x: np.stack([oracle.fk(desc, xi)["r"] for xi in x]))
    except Exception as e:
        print('sample fail', e); raise
    base = None
    for mu0, thr0 in [(0.1,3.0),(0.1,1.0),(0.01,1.0),(0.01,0.3),(1.0,3.0),(0.01,3.0),(1e-3,0.3),(1e-3,0.1),(0.1,0.3)]:
        desc.qp_mu0, desc.qp_thr0 = mu0, thr0
        out = oracle.solve_batch(desc, batch["x0"], batch["target"], batch["body_params"])
        it = out["stats"][:,0]
        if base is None: base = out
        err = np.abs(out["X"]-base["X"]).max()
        print(f"{name:24s} mu0 {mu0:6.3f} thr0 {thr0:4.1f}: iters mean {it.mean():5.2f} max {it.max():3.0f} status {np.bincount(out['status'],minlength=4)} dX vs base {err:.2e}")
