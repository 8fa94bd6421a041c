"""Per-phase cycle breakdown of the solve kernel (needs a library built with -DUB_PROFILE=1: tools/build_variant.py profile
-DUB_PROFILE=1, UB_LIBRARY=variants/profile/libupright_b200.so).  Third argument "persistent": counters of the product
grid (work queue, alignment) instead of the static test grid."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_thing_demo"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
desc, meta = workload.load(name)
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))
b = workload.sample_batch(name, desc, meta, B, 1234, ee)
dev = lambda a: None if a is None else torch.tensor(a, dtype=torch.float32, device="cuda")
x0, tg, bp = dev(b["x0"]), dev(b["target"]), dev(b["body_params"])
for _ in range(2):
    out = mpc.solve_device(x0, tg, bp)
torch.cuda.synchronize()
print("normal ms", mpc.last_solve_ms())
BASE = 10 if len(sys.argv) > 3 and sys.argv[3] == "persistent" else 0   # persistent: counters of the product grid
mpc.set_option("stop_after", BASE + 9)
out = mpc.solve_device(x0, tg, bp)
torch.cuda.synchronize()
st = out["stats"].double().cpu().numpy()
names = ["qp_iters", "A:gradient", "A:build", "A:dynamics", "A:cholesky", "A total (factor+predict bwd)", "B fwd predictor + mu_aff", "C+D corrector bwd+fwd"]
tot = st[:, 5:].sum(1)
print("profile-mode ms", mpc.last_solve_ms(), "mean iters", st[:, 0].mean(), "max iters", st[:, 0].max())
for i in range(1, 8):
    print(f"  {names[i]:14s} mean {st[:, i].mean() / 1e6:8.3f} Mcyc  ({100 * st[:, i].mean() / tot.mean():5.1f} %)  per-iter {st[:, i].mean() / st[:, 0].mean() / 1e3:8.1f} kcyc")
print("  total per warp %.2f Mcyc" % (tot.mean() / 1e6))

mpc.set_option("stop_after", BASE + 8)
out = mpc.solve_device(x0, tg, bp)
torch.cuda.synchronize()
st = out["stats"].double().cpu().numpy()
tot = st[:, 3].mean()
print("whole-solve profile (Mcyc per warp): total %.2f | linearise %.2f (%.0f %%) | interior point %.2f (%.0f %%) | line search %.2f (%.0f %%) | "
      "init (Df + base performance) %.2f (%.0f %%)" % (tot / 1e6, st[:, 1].mean() / 1e6, 100 * st[:, 1].mean() / tot, st[:, 4].mean() / 1e6,
                                                     100 * st[:, 4].mean() / tot, st[:, 2].mean() / 1e6, 100 * st[:, 2].mean() / tot,
                                                     st[:, 5].mean() / 1e6, 100 * st[:, 5].mean() / tot))
print("  waiting at the alignment meetings %.2f Mcyc (%.0f %%) | set-up of the interior-point iteration %.2f Mcyc (%.0f %%) | per iteration %.3f Mcyc | "
      "kernel ms %.3f" % (st[:, 6].mean() / 1e6, 100 * st[:, 6].mean() / tot, st[:, 7].mean() / 1e6, 100 * st[:, 7].mean() / tot,
                          st[:, 4].mean() / 1e6 / max(1e-9, st[:, 0].mean()), mpc.last_solve_ms()))
