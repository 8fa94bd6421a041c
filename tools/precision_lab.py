"""Precision study on the CPU (no GPU needed): the oracle's interior-point QP step in three arithmetics —
fp64, fp32 throughout (the arithmetic of the fp32 kernels), and fp32 Newton matrices / factors / directions with the
iterate, slack records and residuals in fp64 (the refinement proposed in DESIGN.md section 9) — over seeded samples
of the BASELINE configurations and the light-body soft-row case of section 8.

    python tools/precision_lab.py [instances per configuration, default 48]
"""
import copy
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from upright_b200 import problem_io, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48


def cases():
    for name in problem_io.FIXTURES:
        desc, meta = workload.load(name)
        yield name, desc, meta
    desc, meta = workload.load("cfg4_thing_obstacles2")
    soft = copy.deepcopy(desc)
    soft.slacks.enabled = 1
    yield "cfg4 + slacks (20 g body, soft rows)", soft, meta


def ranges(desc):
    nq, nx = desc.nq, 3 * desc.nq
    nu = nq + (desc.nf * desc.nc if desc.balancing_enabled else 0)
    rx = np.array(desc.state_ub[:nx]) - np.array(desc.state_lb[:nx])
    ru = np.concatenate([np.array(desc.input_ub[:nq]) - np.array(desc.input_lb[:nq]),
                         np.full(nu - nq, desc.force_ub - desc.force_lb)])
    return rx, ru


print(f"{'configuration':40s} {'mode':22s} {'conv':>5s} {'fail':>5s} {'iters':>6s} {'median err':>11s} {'p95 err':>9s} {'max err':>9s}")
for name, desc, meta in cases():
    base = name.split(" ")[0] if name.startswith("cfg4 +") else name
    fixture = "cfg4_thing_obstacles2" if name.startswith("cfg4 +") else base
    nu = oracle.dims(desc)["nu"]
    ee = lambda x: np.array([oracle.fk(desc, xi)["r"] for xi in x])  # noqa: E731
    mg = (lambda x: np.array([oracle.linearize(desc, xi, np.zeros(nu))["hobs"] for xi in x])) if desc.obstacles_enabled else None
    b = workload.sample_batch(fixture, desc, meta, B, 1234, ee, margin_fn=mg)
    rx, ru = ranges(desc)
    MODES = ((0, 'fp64'), (1, 'fp32'), (2, 'fp32 factors + fp64 it.'), (10, 'fp64, +1 iteration'), (11, 'fp32, +1 iteration'),
             (12, 'mixed, +1 iteration'), (20, 'fp64, +2 iterations'), (22, 'mixed, +2 iterations'),
             (100, 'fp64, mu <= 1.2 target'), (101, 'fp32, mu <= 1.2 target'), (200, 'fp64, mu <= 1.05 target'),
             (201, 'fp32, mu <= 1.05 target'), (202, 'mixed, mu <= 1.05 target'))   # compared with fp64 of the same variant
    res = {m: [] for m, _ in MODES}
    for i in range(B):
        X = np.tile(b["x0"][i], (desc.N + 1, 1))
        U = np.zeros((desc.N, nu))
        bp = None if b["body_params"] is None else b["body_params"][i]
        for m, _ in MODES:
            res[m].append(oracle.qp_step_precision(desc, b["target"][i], X, U, m, bp))
    for m, label in MODES:
        conv = sum(r["converged"] for r in res[m])
        fail = sum(r["failed"] > 0 for r in res[m])
        iters = np.mean([r["iters"] for r in res[m]])
        errs = []
        for r, r0 in zip(res[m], res[10 * (m // 10)]):
            if r["failed"] or not r0["converged"] or not np.isfinite(r["dX"]).all():
                continue
            errs.append(max((np.abs(r["dX"] - r0["dX"]) / rx).max(), (np.abs(r["dU"] - r0["dU"]) / ru).max()))
        errs = np.array(errs) if errs else np.array([np.nan])
        print(f"{name:40s} {label:22s} {conv:5d} {fail:5d} {iters:6.1f} {np.median(errs):11.2e} {np.percentile(errs, 95):9.2e} {errs.max():9.2e}")
