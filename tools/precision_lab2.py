"""Precision study of the reduced (force-eliminated) stage — oracle/reduced_lab.h — against the fp64 dense oracle step.
    python tools/precision_lab2.py [instances] [modes comma list] [cfg filter]
"""
import copy
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import oracle  # noqa: E402
from upright_b200 import problem_io, workload  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
LABEL = {0: "dense fp64", 1: "dense fp32", 2: "dense mixed", 3: "reduced fp64", 4: "reduced F32 R64 S64 data32",
         5: "reduced F32 R64 S64 data64", 6: "reduced F32 R64 S32 data32", 7: "reduced all fp32"}
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3, 4, 5, 6, 14]
filt = sys.argv[3] if len(sys.argv) > 3 else ""


def cases():
    for name in problem_io.FIXTURES:
        desc, meta = workload.load(name)
        yield name, name, desc, meta
    desc, meta = workload.load("cfg4_thing_obstacles2")
    soft = copy.deepcopy(desc)
    soft.slacks.enabled = 1
    yield "cfg4+slacks(20g)", "cfg4_thing_obstacles2", soft, meta


def ranges(desc):
    nq, nx = desc.nq, 3 * desc.nq
    nu = nq + (desc.nf * desc.nc if desc.balancing_enabled else 0)
    rx = np.array(desc.state_ub[:nx]) - np.array(desc.state_lb[:nx])
    ru = np.concatenate([np.array(desc.input_ub[:nq]) - np.array(desc.input_lb[:nq]), np.full(nu - nq, desc.force_ub - desc.force_lb)])
    return rx, ru


print(f"{'configuration':22s} {'mode':30s} {'conv':>5s} {'fail':>5s} {'iters':>6s} {'median':>9s} {'p95':>9s} {'max':>9s}")
for name, fixture, desc, meta in cases():
    if filt and filt not in name:
        continue
    nu = oracle.dims(desc)["nu"]
    ee = lambda x: np.array([oracle.fk(desc, xi)["r"] for xi in x])  # noqa: E731
    mg = (lambda x: np.array([oracle.linearize(desc, xi, np.zeros(nu))["hobs"] for xi in x])) if desc.obstacles_enabled else None
    b = workload.sample_batch(fixture, desc, meta, B, 1234, ee, margin_fn=mg)
    rx, ru = ranges(desc)
    res = {m: [] for m in [0] + modes}
    for i in range(B):
        X = np.tile(b["x0"][i], (desc.N + 1, 1))
        U = np.zeros((desc.N, nu))
        bp = None if b["body_params"] is None else b["body_params"][i]
        for m in res:
            res[m].append(oracle.qp_step_precision(desc, b["target"][i], X, U, m, bp))
    for m in modes:
        conv = sum(r["converged"] for r in res[m])
        fail = sum(r["failed"] > 0 for r in res[m])
        iters = np.mean([r["iters"] for r in res[m]])
        errs = []
        for r, r0 in zip(res[m], res[0]):
            if r["failed"] or not r0["converged"] or not r["converged"] or not np.isfinite(r["dX"]).all():
                continue
            errs.append(max((np.abs(r["dX"] - r0["dX"]) / rx).max(), (np.abs(r["dU"] - r0["dU"]) / ru).max()))
        errs = np.array(errs) if errs else np.array([np.nan])
        lab = LABEL[m % 10] + (f" +{(m // 10) % 10}it" if (m // 10) % 10 else "")
        print(f"{name:22s} {lab:30s} {conv:5d} {fail:5d} {iters:6.1f} {np.median(errs):9.2e} {np.percentile(errs, 95):9.2e} {errs.max():9.2e}", flush=True)
