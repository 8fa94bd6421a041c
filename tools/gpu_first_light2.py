"""First light of the reduced-stage kernels: fp64 and mixed kernels against the oracle per configuration, and — to
localise a discrepancy — the QP iterate after ONE interior-point iteration (qp_iter_max = 1, stop_after = 2) split
into its jerk / force / state parts per stage."""
import copy
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from upright_b200 import problem_io, workload  # noqa: E402
from upright_b200.engine import BatchedMPC  # noqa: E402
np.set_printoptions(precision=3, suppress=False, linewidth=220)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(problem_io.FIXTURES)
Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 16


def ranges(desc):
    nq, nx, nu = desc.nq, desc.nx, desc.nu
    rx = np.array(desc.state_ub[:nx]) - np.array(desc.state_lb[:nx])
    ru = np.concatenate([np.array(desc.input_ub[:nq]) - np.array(desc.input_lb[:nq]), np.full(nu - nq, desc.force_ub - desc.force_lb)])
    return rx, ru


for name in names:
    desc, meta = workload.load(name)
    nq, nx, nu, N = desc.nq, desc.nx, desc.nu, desc.N
    nz = nx + nu
    probe = BatchedMPC(desc, "f64")
    ee = lambda x: probe.eval("end_effector_position", x, np.zeros((x.shape[0], nu)))
    mg = (lambda x: probe.eval("obstacle_avoidance", x, np.zeros((x.shape[0], nu)))) if desc.obstacles_enabled else None
    b = workload.sample_batch(name, desc, meta, Bn, 7, ee, margin_fn=mg)
    rx, ru = ranges(desc)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    for prec in ("f64", "f32"):
        try:
            m = BatchedMPC(desc, prec)
            out = m.solve(b["x0"], b["target"], b["body_params"], want_gains=(prec == "f64"))
        except Exception as e:  # noqa: BLE001
            print(name, prec, "FAILED", e)
            continue
        ex = (np.abs(out["X"] - ref["X"]) / rx).reshape(Bn, -1).max(1)
        eu = (np.abs(out["U"] - ref["U"]) / ru).reshape(Bn, -1).max(1)
        good = (ref["status"] == 0) & (out["status"] == 0)
        e = np.maximum(ex, eu)[good] if good.any() else np.array([np.nan])
        print(f"{name:24s} {prec} status {np.bincount(out['status'], minlength=4)} ref {np.bincount(ref['status'], minlength=4)} "
              f"err max {e.max():.2e} med {np.median(e):.2e} | iters {out['stats'][:8, 0].astype(int)} ref {ref['stats'][:8, 0].astype(int)} "
              f"nanreason {out['stats'][out['status'] == 3, 3][:4]}", flush=True)
        if prec == "f64":
            refk = oracle.solve_batch(desc, b["x0"][:4], b["target"][:4], None if b["body_params"] is None else b["body_params"][:4], want_gains=True)
            kerr = np.abs(out["K"][:4] - refk["K"]).max() / max(1.0, np.abs(refk["K"]).max())
            print(f"{'':24s} gains rel err {kerr:.2e} (jerk rows {np.abs(out['K'][:4, :, :nq] - refk['K'][:, :, :nq]).max():.2e}, force rows {np.abs(out['K'][:4, :, nq:] - refk['K'][:, :, nq:]).max():.2e})")
    # one interior-point iteration
    d1 = copy.deepcopy(desc)
    d1.qp_iter_max = 1
    for prec in ("f64", "f32"):
        m1 = BatchedMPC(d1, prec)
        dt = m1.torch_dtype
        m1.set_option("stop_after", 2)
        dev = lambda a: None if a is None else torch.tensor(a, dtype=dt, device="cuda")
        nb = min(Bn, 4)
        m1.solve_device(dev(b["x0"][:nb]), dev(b["target"][:nb]), dev(None if b["body_params"] is None else b["body_params"][:nb]))
        torch.cuda.synchronize()
        Z = m1.workspace_block(nb, "Z", (N + 1) * nz).reshape(nb, N + 1, nz)
        for i in range(nb):
            X0 = np.tile(b["x0"][i], (N + 1, 1))
            dX, dU, info = oracle.qp_step(d1, b["target"][i], X0, np.zeros((N, nu)), None if b["body_params"] is None else b["body_params"][i])
            ej = np.abs(Z[i, :N, :nq] - dU[:, :nq]).max(axis=1)
            ef = np.abs(Z[i, :N, nq:nu] - dU[:, nq:]).max(axis=1) if nu > nq else np.zeros(N)
            exx = np.abs(Z[i, :, nu:] - dX).max(axis=1)
            print(f"   1-iter {prec} inst {i}: |dz| ref max j {np.abs(dU[:, :nq]).max():.2e} f {np.abs(dU[:, nq:]).max() if nu > nq else 0:.2e} x {np.abs(dX).max():.2e} "
                  f"| err j {ej.max():.2e} (stage {ej.argmax()}) f {ef.max():.2e} (stage {ef.argmax()}) x {exx.max():.2e} (stage {exx.argmax()})", flush=True)
            if i == 0 and max(ej.max(), ef.max(), exx.max()) > 1e-6 * max(1.0, np.abs(dU).max()) and prec == "f64":
                print("      per-stage err j", ej, "\n      per-stage err f", ef, "\n      per-stage err x", exx)
