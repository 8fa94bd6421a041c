"""First-light check on the GPU box: linearisation blocks, QP step and full
solve of the CUDA path against the oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from upright_b200 import problem_io  # noqa: E402
from upright_b200.engine import BatchedMPC  # noqa: E402

np.set_printoptions(precision=5, suppress=True, linewidth=200)
names = sys.argv[1:] or ["cfg2_thing_demo"]
for name in names:
    desc, meta = problem_io.load_fixture(name)
    N, nx, nu, nq = desc.N, desc.nx, desc.nu, desc.nq
    rng = np.random.default_rng(0)
    Bn = 4
    x0 = np.tile(meta["x0"], (Bn, 1))
    x0[1:, :nq] += rng.uniform(-0.2, 0.2, (Bn - 1, nq))
    wp = np.tile(meta["waypoint"], (Bn, 1))
    wp[1:] = rng.uniform(-0.3, 0.3, (Bn - 1, 3))
    r0 = np.array([oracle.fk(desc, x)["r"] for x in x0])
    target = np.repeat((r0 + wp)[:, None, :], N + 1, axis=1)
    ref = oracle.solve_batch(desc, x0, target)
    for prec in ("f64", "f32"):
        mpc = BatchedMPC(desc, prec)
        dt = mpc.torch_dtype
        # 1. linearisation
        mpc.set_option("stop_after", 1)
        out = mpc.solve_device(torch.tensor(x0, dtype=dt, device="cuda"), torch.tensor(target, dtype=dt, device="cuda"))
        torch.cuda.synchronize()
        ws, L = mpc.workspace_view(Bn)
        ws = ws.cpu().numpy().astype(np.float64)
        b = 1
        X0 = np.tile(x0[b], (N + 1, 1))
        lin = oracle.linearize(desc, X0[3], np.zeros(nu))
        neq = mpc.n_eq
        if neq:
            nz = nx + nu
            rows = ws[b, L["LCT"] + 3 * neq * nz: L["LCT"] + 4 * neq * nz].reshape(neq, nz)
            g = ws[b, L["LG"] + 3 * neq: L["LG"] + 4 * neq]
            print(name, prec, "lin: |C-C_or|", np.abs(rows[:, nu:] - lin["C"]).max(), "|g-g_or|", np.abs(g - lin["g"]).max())
        Jp = ws[b, L["LJP"] + 3 * 3 * nq: L["LJP"] + 4 * 3 * nq].reshape(3, nq)
        print(name, prec, "lin: |Jp|", np.abs(Jp - lin["Jp"]).max(), "|r|", np.abs(ws[b, L["LR"] + 9: L["LR"] + 12] - lin["r"]).max())
        # 2. QP step
        mpc.set_option("stop_after", 2)
        out = mpc.solve_device(torch.tensor(x0, dtype=dt, device="cuda"), torch.tensor(target, dtype=dt, device="cuda"))
        torch.cuda.synchronize()
        ws, L = mpc.workspace_view(Bn)
        ws = ws.cpu().numpy().astype(np.float64)
        for b in range(Bn):
            X0 = np.tile(x0[b], (N + 1, 1))
            dX, dU, info = oracle.qp_step(desc, target[b], X0, np.zeros((N, nu)))
            Z = ws[b, L["Z"]: L["Z"] + (N + 1) * (nx + nu)].reshape(N + 1, nx + nu)
            print(name, prec, b, "qp: |dX|", np.abs(Z[:, nu:] - dX).max(), "|dU|", np.abs(Z[:N, :nu] - dU).max(), info)
        # 3. full solve (host path)
        mpc.set_option("stop_after", 0)
        res = mpc.solve(x0, target)
        print(name, prec, "status", res["status"], "ref", ref["status"])
        print(name, prec, "stats gpu", res["stats"][:, :4], "\n  stats ref", ref["stats"][:, :4])
        print(name, prec, "solve: |X|", np.abs(res["X"] - ref["X"]).max(), "|U|", np.abs(res["U"] - ref["U"]).max(),
              "ms", mpc.last_solve_ms())
