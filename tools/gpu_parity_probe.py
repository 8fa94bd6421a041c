"""Error distribution of the CUDA path against the oracle on the random workload."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from upright_b200 import workload  # noqa: E402
from upright_b200.engine import BatchedMPC  # noqa: E402
np.set_printoptions(precision=4, suppress=True, linewidth=220)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_thing_demo"
Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 64
desc, meta = workload.load(name)
nq, nx, nu, N = desc.nq, desc.nx, desc.nu, desc.N
m64 = BatchedMPC(desc, "f64")
m32 = BatchedMPC(desc, "f32")
ee = lambda x: m64.eval("end_effector_position", x, np.zeros((x.shape[0], nu)))
b = workload.sample_batch(name, desc, meta, Bn, 7, ee)
ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
rx = np.concatenate([np.array(desc.state_ub[:nx]) - np.array(desc.state_lb[:nx])])
ru = np.concatenate([np.array(desc.input_ub[:nq]) - np.array(desc.input_lb[:nq]), np.full(nu - nq, desc.force_ub - desc.force_lb)])
for tag, m in (("f64", m64), ("f32", m32)):
    out = m.solve(b["x0"], b["target"], b["body_params"])
    ex = np.abs(out["X"] - ref["X"]) / rx
    eu = np.abs(out["U"] - ref["U"]) / ru
    exi, eui = ex.reshape(Bn, -1).max(1), eu.reshape(Bn, -1).max(1)
    print(tag, "status gpu", np.bincount(out["status"], minlength=4), "ref", np.bincount(ref["status"], minlength=4))
    print(tag, "scaled err X: max %.2e median %.2e | U: max %.2e median %.2e" % (exi.max(), np.median(exi), eui.max(), np.median(eui)))
    print(tag, "iters gpu", out["stats"][:, 0].astype(int)[:24], "\n    ref  ", ref["stats"][:, 0].astype(int)[:24])
    print(tag, "alpha gpu", out["stats"][:, 3][:24], "\n    ref  ", ref["stats"][:, 3][:24])
    worst = np.argsort(-np.maximum(exi, eui))[:5]
    for w in worst:
        k, i = np.unravel_index(np.argmax(ex[w]), ex[w].shape)
        print("   worst", w, "exi %.2e eui %.2e" % (exi[w], eui[w]), "it", out["stats"][w, 0], ref["stats"][w, 0], "alpha", out["stats"][w, 3], ref["stats"][w, 3],
              "st", out["status"][w], ref["status"][w], "at knot", k, "comp", i, "cost", out["stats"][w, 1], ref["stats"][w, 1], "viol", out["stats"][w, 2], ref["stats"][w, 2])
