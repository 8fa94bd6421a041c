import sys, copy, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle
from upright_b200 import problem_io, geometry as geo, workload
from upright_b200.engine import BatchedMPC
base, meta = problem_io.load_fixture("cfg2_thing_demo")
for mode in ("plain", "fixed"):
    desc = copy.deepcopy(base)
    desc.ia_constraint_enabled = 1; desc.ia_alpha = 0.05
    desc.ia_normal[:] = [0, 0, 1.0]; desc.ia_span[:] = geo.plane_span([0, 0, 1]).reshape(6); desc.ia_com[:] = [0.02, -0.01, 0.15]
    desc.ia_align_with_fixed_vector = int(mode == "fixed")
    m = BatchedMPC(desc, "f64")
    ee = lambda x: m.eval("end_effector_position", x, np.zeros((x.shape[0], m.nu)))
    b = workload.sample_batch("cfg2_thing_demo", desc, meta, 8, 17, ee)
    m.set_option("stop_after", 1)
    dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
    out = m.solve_device(dev(b["x0"]), dev(b["target"]), dev(b["body_params"]))
    torch.cuda.synchronize()
    ws, L = m.workspace_view(8)
    ws = ws.cpu().numpy()
    nobs, nx, nu = 5, 27, 13
    X = np.tile(b["x0"][0], (21, 1)); U = np.zeros((20, 13))
    q = oracle.qp_dump(desc, b["target"][0], X, U, b["body_params"][0])
    for k in (1, 5):
        J = ws[0, L["LJO"] + k * nobs * nx: L["LJO"] + (k + 1) * nobs * nx].reshape(5, nx)
        h = ws[0, L["LHO"] + k * nobs: L["LHO"] + (k + 1) * nobs]
        A5, c5 = q[k]["A"][-5:, nu:], q[k]["c"][-5:]
        print(mode, k, "dh", np.abs(h - c5).max(), "dJ", np.abs(J - A5).max(), "|J|", np.abs(A5).max())
    m.set_option("stop_after", 0)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    o = m.solve(b["x0"], b["target"], b["body_params"])
    print(mode, "solve diff", np.abs(o["X"] - ref["X"]).max(), "iters", o["stats"][:, 0], ref["stats"][:, 0])
