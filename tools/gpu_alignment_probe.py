"""How much does instruction-stream alignment between the warps of an SM matter?  Same total work (same number of
interior-point iterations), once with instances that all take 5 iterations (every warp of the persistent grid stays in
step: same code at the same time) and once with a 4 / 5 / 6 mixture (warps drift apart after the first wave)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upright_b200 import workload
from upright_b200.engine import BatchedMPC
name = "cfg2_thing_demo"
desc, meta = workload.load(name)
mpc = BatchedMPC(desc, "f32")
ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))  # noqa: E731
pool = workload.sample_batch(name, desc, meta, 65536, 99, ee)
dev = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")  # noqa: E731
x0, tg, bp = dev(pool["x0"]), dev(pool["target"]), dev(pool["body_params"])
out = mpc.solve_device(x0, tg, bp)
torch.cuda.synchronize()
it = out["stats"][:, 0].long().cpu().numpy()
print("pool iterations:", {int(k): int((it == k).sum()) for k in np.unique(it)})
slots = 148 * 16
W = 8
n = W * slots
i5 = np.where(it == 5)[0]
i4, i6 = np.where(it == 4)[0], np.where(it == 6)[0]
m = min(len(i4), len(i6), n // 3)
rng = np.random.default_rng(0)
setA = i5[:n]
setM = np.concatenate((i4[:m], i6[:m], i5[: n - 2 * m]))
rng.shuffle(setM)
assert len(setA) == n and len(setM) == n, (len(setA), len(setM))
print(f"batch {n} = {W} waves of {slots}; mixture: {m} x 4, {m} x 6, {n - 2 * m} x 5 iterations (same total)")
for label, idx in (("all 5 iterations (aligned)", setA), ("4/5/6 mixture (drifting)", setM), ("all 5 iterations (aligned)", setA),
                   ("4/5/6 mixture (drifting)", setM)):
    ii = torch.tensor(idx, device="cuda")
    a, b, c = x0[ii].contiguous(), tg[ii].contiguous(), bp[ii].contiguous()
    ts = []
    for _ in range(3):
        o = mpc.solve_device(a, b, c)
        torch.cuda.synchronize()
        ts.append(mpc.last_solve_ms())
    tot = int(o["stats"][:, 0].sum().item())
    print(f"{label:30s}: {min(ts):8.3f} ms, total iterations {tot}, {min(ts) / tot * 1e6:7.2f} ns per iteration-instance")
