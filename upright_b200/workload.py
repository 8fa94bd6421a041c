"""Seeded synthetic batches for the BASELINE configurations:
random start states around the shipped home configuration, random goal
offsets inside the box spanned by the shipped waypoints, per-instance
inertial parameters with fixed contact topology."""
from __future__ import annotations

import numpy as np

from . import problem_io

# goal offset box per configuration [lo, hi] (m), relative to the start EE position.
# cfg2/3/5: box spanning the shipped waypoints (thing_demo.yaml:55,
# upright_robust/scripts/planning_sim_loop.py:449).  cfg1/cfg4 run with hard
# terminal constraints (slacks disabled), so goals stay within what a 2 s horizon
# can reach under the state limits.
GOAL_BOX = {
    "cfg1_ur10_demo": ([-0.25, -0.25, -0.2], [0.25, 0.25, 0.2]),
    "cfg2_thing_demo": ([-2.0, -2.0, -0.25], [2.0, 1.0, 0.25]),
    "cfg3_thing_box_arch": ([-2.0, -2.0, -0.25], [2.0, 1.0, 0.25]),
    "cfg4_thing_obstacles2": ([-0.15, -0.15, -0.1], [0.15, 0.15, 0.1]),
    "cfg5_thing_robust8": ([-2.0, -2.0, -0.25], [2.0, 1.0, 0.25]),
}
BASELINE_BATCH = {
    "cfg1_ur10_demo": 1,
    "cfg2_thing_demo": 4096,
    "cfg3_thing_box_arch": 4096,
    "cfg4_thing_obstacles2": 16384,
    "cfg5_thing_robust8": 8192,
}


def sample_batch(name, desc, meta, B, seed, ee_position_fn, vary_bodies=True, level_tray=None, margin_fn=None):
    """-> dict(x0 [B,nx], target [B,N+1,3], body_params [B,nb,10] or None).

    `ee_position_fn(x [M,nx]) -> [M,3]` evaluates the tool position (GPU probe
    or the oracle).  `level_tray=True` perturbs only joints that keep the tray
    level (base x, y, yaw and shoulder pan), which hard-constrained
    configurations need to be feasible at the first knot.
    """
    rng = np.random.default_rng(seed)
    nq, nx, N = desc.nq, 3 * desc.nq, desc.N
    if level_tray is None:
        level_tray = not bool(desc.slacks.enabled)
    x0 = np.tile(np.asarray(meta["x0"], dtype=float), (B, 1))
    dq = rng.uniform(-0.3, 0.3, (B, nq))
    if nq == 9:
        dq[:, 0:2] = rng.uniform(-1.0, 1.0, (B, 2))
        dq[:, 2] = rng.uniform(-0.5, 0.5, B)
        free = [0, 1, 2, 3]
    else:
        free = [0]
    if desc.obstacles_enabled and nq == 9:
        dq[:, 0:2] *= 0.25
        dq[:, 2] *= 0.5
    if level_tray:
        mask = np.zeros(nq, dtype=bool)
        mask[free] = True
        dq[:, ~mask] = 0.0
    x0[:, :nq] += dq
    if desc.obstacles_enabled and margin_fn is not None:
        home = np.asarray(meta["x0"], dtype=float)
        for _ in range(20):
            bad = margin_fn(x0).min(axis=1) < 0.02
            if not bad.any():
                break
            scale = rng.uniform(0.0, 1.0, (int(bad.sum()), 1))
            x0[bad, :nq] = home[:nq] + scale * (x0[bad, :nq] - home[:nq])
    lo, hi = GOAL_BOX.get(name, ([-0.5, -0.5, -0.2], [0.5, 0.5, 0.2]))
    goal = ee_position_fn(x0) + rng.uniform(lo, hi, (B, 3))
    target = np.repeat(goal[:, None, :], N + 1, axis=1)
    body = None
    if vary_bodies and desc.nb > 0 and desc.balancing_enabled:
        base = np.array([[desc.body_params[b][j] for j in range(10)] for b in range(desc.nb)])
        body = np.tile(base, (B, 1, 1))
        mscale = rng.uniform(0.8, 1.2, (B, 1, 1))
        body = body * mscale  # mass, m*com and inertia all scale with mass
        dcom = rng.uniform(-0.01, 0.01, (B, desc.nb, 3))
        body[:, :, 1:4] += body[:, :, 0:1] * dcom
    return dict(x0=x0, target=target, body_params=body)


def load(name):
    return problem_io.load_fixture(name)


def algorithmic_bytes_per_solve(desc, with_gains=False):
    """compulsory fp32 I/O of one solve."""
    nx, nu, N = 3 * desc.nq, desc.nq + desc.nf * desc.nc * bool(desc.balancing_enabled), desc.N
    b_in = 4 * (nx + 8 + 10 * desc.nb + 16 * desc.nc)
    b_out = 4 * ((N + 1) * nx + N * nu)
    return b_in + b_out + (4 * N * nu * nx if with_gains else 0)


def algorithmic_flops_per_solve(desc, ipm_iters, sqp_iters=1):
    """SURVEY.md §8(d) counting convention:
    F = S [ (N+1) F_lin + I N (F_ric + F_con) + F_ls ]."""
    nq = desc.nq
    bal = bool(desc.balancing_enabled)
    nx, nu, N = 3 * nq, nq + (desc.nf * desc.nc if bal else 0), desc.N
    f_ric = 7.0 / 3 * nx**3 + 4 * nx**2 * nu + 2 * nx * nu**2 + nu**3 / 3.0
    nfc = nu - nq
    rows = [1] * (nu + nx)  # box rows
    if bal:
        rows += [nx + (desc.nf * 4 if desc.nc else 0)] * (6 * desc.nb)  # object-dynamics rows
        if desc.nf == 3:
            rows += [3] * (5 * desc.nc)
    if desc.obstacles_enabled:
        rows += [nq] * desc.n_pairs
    f_con = 2.0 * sum(r * r for r in rows)
    f_lin = 2.0e4
    f_ls = 3 * (N + 1) * 2.0e3
    del nfc
    return sqp_iters * ((N + 1) * f_lin + ipm_iters * N * (f_ric + f_con) + f_ls)
