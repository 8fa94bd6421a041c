"""The oracle's QP solver, checked without trusting it:
(1) KKT conditions of the returned step verified by an independent dense numpy
    computation (stationarity in the null space of the dynamics, residual
    definitions, complementarity at the central-path target);
(2) the interior-point method and the semismooth-Newton method (two unrelated
    algorithms) converge to the same minimiser as mu -> 0."""
import numpy as np
import pytest

from upright_b200 import problem_io


def _setup(name):
    desc, meta = problem_io.load_fixture(name)
    N, nx, nu = desc.N, desc.nx, desc.nu
    target = np.tile(meta["r_ee0"] + meta["waypoint"], (N + 1, 1))
    rng = np.random.default_rng(11)
    x0 = np.array(meta["x0"], dtype=float)
    if desc.slacks.enabled:
        x0[: desc.nq] += 0.1 * rng.standard_normal(desc.nq)
    X = np.tile(x0, (N + 1, 1))
    U = np.zeros((N, nu))
    return desc, target, X, U


def _dynamics(desc):
    nq, nx, nu, dt = desc.nq, desc.nx, desc.nu, desc.dt
    A, Bm = np.eye(nx), np.zeros((nx, nu))
    I = np.eye(nq)
    A[:nq, nq:2 * nq] = dt * I
    A[:nq, 2 * nq:] = 0.5 * dt * dt * I
    A[nq:2 * nq, 2 * nq:] = dt * I
    Bm[:nq, :nq] = dt**3 / 6 * I
    Bm[nq:2 * nq, :nq] = 0.5 * dt * dt * I
    Bm[2 * nq:, :nq] = dt * I
    return A, Bm


def _nullspace_gradient(desc, stages, dX, dU):
    """Gradient of the soft-penalised objective w.r.t. the inputs after
    eliminating the states through the dynamics (adjoint recursion)."""
    N, nx, nu = desc.N, desc.nx, desc.nu
    A, Bm = _dynamics(desc)
    grads = []
    for k, s in enumerate(stages):
        z = np.concatenate([dU[k], dX[k]]) if k < N else dX[k]
        g = s["H"] @ z + s["g"]
        val = s["A"] @ z + s["c"]
        resid = val - np.clip(val, s["lb"], s["ub"])
        g = g + s["A"].T @ (s["rho"] * resid)
        grads.append(g)
    lam = grads[N]
    out = []
    for k in range(N - 1, -1, -1):
        gu, gx = grads[k][:nu], grads[k][nu:]
        out.append(gu + Bm.T @ lam)
        lam = gx + A.T @ lam
    return np.concatenate(out[::-1])


@pytest.mark.parametrize("name", ["cfg2_thing_demo", "cfg5_thing_robust8"])
def test_ssn_solution_is_stationary(oracle_lib, name):
    """Soft rows only: the SSN result must zero the reduced gradient of the
    piecewise-quadratic objective (convex => global minimiser)."""
    desc, target, X, U = _setup(name)
    if name == "cfg5_thing_robust8":
        desc.slacks.input_box = 1  # make every row soft for this check
    desc.qp_method, desc.qp_iter_max = 1, 400
    stages = oracle_lib.qp_dump(desc, target, X, U)
    assert not any(s["hard"].any() for s in stages)
    dX, dU, info = oracle_lib.qp_step(desc, target, X, U)
    assert info["converged"]
    A, Bm = _dynamics(desc)
    for k in range(desc.N):  # dynamics feasibility of the step
        assert np.allclose(dX[k + 1], A @ dX[k] + Bm @ dU[k] + stages[k]["b"], atol=1e-10)
    g = _nullspace_gradient(desc, stages, dX, dU)
    scale = max(1.0, max(np.abs(s["g"]).max() for s in stages))
    assert np.abs(g).max() < 1e-7 * scale


@pytest.mark.parametrize("name", ["cfg2_thing_demo", "cfg3_thing_box_arch", "cfg5_thing_robust8"])
def test_ipm_converges_to_ssn_minimiser(oracle_lib, name):
    desc, target, X, U = _setup(name)
    soft = dict(desc=desc)
    del soft
    desc.slacks.input_box = 1
    desc.qp_method, desc.qp_iter_max = 1, 600
    dXs, dUs, info_s = oracle_lib.qp_step(desc, target, X, U)
    assert info_s["converged"]
    desc.qp_method, desc.qp_iter_max = 0, 80
    errs = []
    for mu in (1e-6, 1e-8, 1e-10):
        desc.qp_mu_target = mu
        dX, dU, info = oracle_lib.qp_step(desc, target, X, U)
        assert info["converged"]
        errs.append(np.abs(dX - dXs).max())
    assert errs[2] < errs[1] < errs[0]
    assert errs[2] < (5e-5 if name == "cfg3_thing_box_arch" else 1e-6)


@pytest.mark.parametrize("name", ["cfg1_ur10_demo", "cfg2_thing_demo", "cfg4_thing_obstacles2"])
def test_ipm_point_satisfies_perturbed_kkt(oracle_lib, name):
    """At the IPM solution: dynamics hold, hard rows are feasible to tolerance,
    and the reduced gradient of  cost + soft penalties - mu * sum log(slack of
    hard inequality sides) + multipliers of hard equalities  vanishes — checked
    through the barrier gradient with lam = mu / t implied by the central path."""
    desc, target, X, U = _setup(name)
    if name == "cfg4_thing_obstacles2":
        target = np.tile(target[0] * 0 + problem_io.load_fixture(name)[1]["r_ee0"] + np.array([0.2, -0.2, 0.0]), (desc.N + 1, 1))
    desc.qp_method = 0
    dX, dU, info = oracle_lib.qp_step(desc, target, X, U)
    assert info["converged"], info
    stages = oracle_lib.qp_dump(desc, target, X, U)
    A, Bm = _dynamics(desc)
    for k in range(desc.N):
        assert np.allclose(dX[k + 1], A @ dX[k] + Bm @ dU[k] + stages[k]["b"], atol=1e-9)
    worst = 0.0
    for k, s in enumerate(stages):
        z = np.concatenate([dU[k], dX[k]]) if k < desc.N else dX[k]
        val = s["A"] @ z + s["c"]
        hard_ineq = s["hard"] & (s["lb"] < s["ub"])
        viol = np.maximum(s["lb"] - val, val - s["ub"])[hard_ineq]
        if viol.size:
            worst = max(worst, viol.max())
    assert worst < 1e-5  # hard inequality rows strictly feasible up to the elastic term


@pytest.mark.parametrize("name", list(problem_io.FIXTURES))
def test_full_solve_reduces_violation_and_is_deterministic(oracle_lib, name):
    desc, meta = problem_io.load_fixture(name)
    target = np.tile(meta["r_ee0"] + (meta["waypoint"] if "cfg4" not in name else np.array([0.2, -0.2, 0.0])), (desc.N + 1, 1))
    a = oracle_lib.solve_batch(desc, meta["x0"], target[None], nthreads=1)
    b = oracle_lib.solve_batch(desc, meta["x0"], target[None], nthreads=2)
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["U"], b["U"])
    assert a["status"][0] == 0
    perf = oracle_lib.performance(desc, target, a["X"][0], a["U"][0])
    assert np.isclose(perf["cost"], a["stats"][0, 1]) and np.isclose(perf["violation"], a["stats"][0, 2])
    assert perf["dyn_sse"] < 1e-16 * max(1.0, a["stats"][0, 3])  # full step => dynamically consistent
    # second SQP iteration (warm) reduces the equality violation
    c = oracle_lib.solve_batch(desc, meta["x0"], target[None], X=a["X"], U=a["U"], warm=True, nthreads=1)
    assert c["stats"][0, 2] < a["stats"][0, 2]


def test_end_effector_box_rows_in_the_oracle():
    """end_effector_box_constraint.h:46-76 — six rows r_d + upper - r >= 0, r - r_d - lower >= 0 at the
    intermediate knots: dims, the performance index and the solve all see them; a box that excludes the
    unconstrained motion changes the solution and keeps the vertical excursion (softly) inside."""
    import copy
    import oracle
    desc, target, X, U = _setup("cfg2_thing_demo")
    free = oracle.solve_batch(desc, X[0], target)
    z_free = np.array([oracle.fk(desc, x)["r"][2] for x in free["X"][0]]) - target[:, 2]
    boxed = copy.deepcopy(desc)
    boxed.ee_box_enabled = 1
    boxed.ee_box_lower[:] = [-5.0, -5.0, -0.02]
    boxed.ee_box_upper[:] = [5.0, 5.0, 0.02]
    assert oracle.dims(boxed)["n_ineq"] == oracle.dims(desc)["n_ineq"] + 6
    # the target sits ~0.25 m above the start: the initial guess violates the lower z row at every knot
    pf = oracle.performance(boxed, target, X, U)
    z0 = oracle.fk(boxed, X[0])["r"][2]
    assert pf["min_margin"] == pytest.approx(z0 - target[0, 2] + 0.02, abs=1e-9) and pf["min_margin"] < -0.1
    assert pf["ineq_sse"] == pytest.approx(desc.dt * (desc.N - 1) * pf["min_margin"] ** 2, rel=1e-9)
    out = oracle.solve_batch(boxed, X[0], target)
    assert out["status"][0] in (0, 1)
    z_box = np.array([oracle.fk(boxed, x)["r"][2] for x in out["X"][0]]) - target[:, 2]
    assert np.abs(out["X"] - free["X"]).max() > 1e-2          # the rows are active
    # pulled into the box as fast as the jerk / acceleration limits allow (soft rows, one SQP step)
    viol = lambda z: float(np.sum(np.minimum(0.0, z[1:-1] + 0.02) ** 2))  # noqa: E731
    assert viol(z_box) < 0.3 * viol(z_free)
    assert z_box[4:-1].min() > -0.05 and z_free[4] < -0.12


def test_inertial_alignment_cost_in_the_oracle():
    """InertialAlignmentCostGaussNewton (inertial_alignment.cpp:90-163): e = S C_we'(a - g)/|g| vanishes for a level
    tray at rest, and enabling the cost tilts/accelerates the tray so that the tangential specific force drops."""
    import copy
    import oracle
    from upright_b200 import geometry as geo
    desc, target, X, U = _setup("cfg2_thing_demo")
    ia = copy.deepcopy(desc)
    ia.ia_cost_enabled = 1
    ia.ia_cost_weight = 10.0
    ia.ia_span[:] = geo.plane_span([0, 0, 1]).reshape(6)
    S = np.array(list(ia.ia_span)).reshape(2, 3)

    def residuals(Xtraj):
        out = []
        for x in Xtraj:
            k = oracle.fk(ia, x)
            C = np.array(k["C"]).reshape(3, 3)
            out.append(S @ (C.T @ (np.array(k["a"]) - np.array(list(ia.gravity)))) / 9.81)
        return np.array(out)

    _, meta = problem_io.load_fixture("cfg2_thing_demo")
    home = np.array(meta["x0"], dtype=float)
    assert np.abs(residuals([home])).max() < 1e-7               # level tray at rest (home angles are rounded)
    # performance index carries dt * 1/2 w e'e
    Xh = np.tile(X[0], (desc.N + 1, 1))
    p0 = oracle.performance(desc, target, Xh, U)["cost"]
    p1 = oracle.performance(ia, target, Xh, U)["cost"]
    e = residuals(Xh[:-1])
    assert p1 - p0 == pytest.approx(desc.dt * 0.5 * 10.0 * np.sum(e * e), rel=1e-9, abs=1e-12)
    # Gauss-Newton model of the QP: the stage gradient / Hessian gain exactly dt w Je'e / dt w Je'Je, checked against
    # central finite differences of the residual itself
    k = 3
    rng = np.random.default_rng(5)
    Xk = Xh.copy()
    Xk[k] += 0.05 * rng.standard_normal(desc.nx)
    q0, q1 = oracle.qp_dump(desc, target, Xk, U)[k], oracle.qp_dump(ia, target, Xk, U)[k]
    nu, nx = desc.nu, desc.nx
    dg, dH = (q1["g"] - q0["g"])[nu:], (q1["H"] - q0["H"])[nu:, nu:]
    assert np.abs((q1["g"] - q0["g"])[:nu]).max() == 0 and np.abs((q1["H"] - q0["H"])[:nu, :]).max() == 0
    h = 1e-6
    Je = np.zeros((2, nx))
    for j in range(nx):
        xp, xm = Xk[k].copy(), Xk[k].copy()
        xp[j] += h
        xm[j] -= h
        Je[:, j] = (residuals([xp])[0] - residuals([xm])[0]) / (2 * h)
    e0 = residuals([Xk[k]])[0]
    assert np.allclose(dg, desc.dt * 10.0 * Je.T @ e0, atol=1e-7)
    assert np.allclose(dH, desc.dt * 10.0 * Je.T @ Je, atol=1e-6) and np.linalg.matrix_rank(dH, tol=1e-9) <= 2
    aligned = oracle.solve_batch(ia, X[0], target)
    assert aligned["status"][0] == 0 and np.isfinite(aligned["X"]).all()


@pytest.mark.parametrize("mode", ["plain", "angular", "fixed"])
def test_inertial_alignment_constraint_rows_in_the_oracle(mode):
    """InertialAlignmentConstraint (inertial_alignment.cpp:7-53): the five rows of the QP of the oracle equal the
    literal formula and its central finite differences, for the three variants."""
    import copy
    import oracle
    from upright_b200 import geometry as geo
    desc, target, X, U = _setup("cfg2_thing_demo")
    ia = copy.deepcopy(desc)
    ia.ia_constraint_enabled = 1
    ia.ia_alpha = 0.3
    ia.ia_normal[:] = [0.0, 0.0, 1.0]
    ia.ia_span[:] = geo.plane_span([0, 0, 1]).reshape(6)
    ia.ia_com[:] = [0.02, -0.01, 0.15]
    ia.ia_use_angular_acceleration = int(mode == "angular")
    ia.ia_align_with_fixed_vector = int(mode == "fixed")
    S, n, com = np.array(list(ia.ia_span)).reshape(2, 3), np.array([0, 0, 1.0]), np.array(list(ia.ia_com))
    g = np.array(list(ia.gravity))

    def rows(x):
        k = oracle.fk(ia, x)
        C, w, al = np.array(k["C"]).reshape(3, 3), np.array(k["w"]), np.array(k["alpha"])
        a = C.T @ (np.array(k["a"]) - g)
        sk = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])  # noqa: E731
        if mode == "angular":
            a = a + (sk(al) + sk(w) @ sk(w)) @ C @ com
        elif mode == "fixed":
            a = C.T @ n
        an, at = n @ a, S @ a
        return np.array([an, 0.3 * an - at[0] - at[1], 0.3 * an - at[0] + at[1], 0.3 * an + at[0] - at[1], 0.3 * an + at[0] + at[1]])

    assert oracle.dims(ia)["n_ineq"] == oracle.dims(desc)["n_ineq"] + 5
    rng = np.random.default_rng(8)
    Xk = np.tile(X[0], (desc.N + 1, 1))
    k = 4
    Xk[k] += 0.1 * rng.standard_normal(desc.nx)
    q0, q1 = oracle.qp_dump(desc, target, Xk, U)[k], oracle.qp_dump(ia, target, Xk, U)[k]
    assert q1["A"].shape[0] == q0["A"].shape[0] + 5
    A5, c5 = q1["A"][-5:], q1["c"][-5:]
    assert np.allclose(c5, rows(Xk[k]), atol=1e-10) and np.abs(A5[:, : desc.nu]).max() == 0
    J = np.zeros((5, desc.nx))
    h = 1e-6
    for j in range(desc.nx):
        xp, xm = Xk[k].copy(), Xk[k].copy()
        xp[j] += h
        xm[j] -= h
        J[:, j] = (rows(xp) - rows(xm)) / (2 * h)
    assert np.allclose(A5[:, desc.nu:], J, atol=2e-6)
    out = oracle.solve_batch(ia, X[0], target)
    assert out["status"][0] in (0, 1) and np.isfinite(out["X"]).all()


def _with_dynamic_obstacle(desc, sphere_slot, state9=None):
    """Copy of `desc` in which the world sphere `sphere_slot` rides on dynamic obstacle 0."""
    import copy
    d = copy.deepcopy(desc)
    assert d.spheres[sphere_slot].link == -1
    d.spheres[sphere_slot].link = -2
    d.n_dynamic_obstacles = 1
    p = np.array(list(d.spheres[sphere_slot].offset))
    return d, (np.concatenate((p, np.zeros(6))) if state9 is None else np.asarray(state9, dtype=float))


def test_dynamic_obstacle_at_rest_equals_static_sphere():
    """A dynamic obstacle (+9 states, obstacle_constraint.h:8-43, system_dynamics.h:28-38) that does not move must
    give exactly the robot trajectory of the static world sphere it replaces; its states stay put."""
    import oracle
    desc, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    slot = [i for i in range(desc.n_spheres) if desc.spheres[i].link == -1][0]
    dyn, xo = _with_dynamic_obstacle(desc, slot)
    assert oracle.dims(dyn)["nx"] == 27 + 9 and oracle.dims(dyn)["n_ineq"] == oracle.dims(desc)["n_ineq"]
    x0 = np.array(meta["x0"], dtype=float)
    target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (desc.N + 1, 1))
    ref = oracle.solve_batch(desc, x0, target)
    out = oracle.solve_batch(dyn, np.concatenate((x0, xo)), target)
    assert out["status"][0] == ref["status"][0] and out["X"].shape == (1, desc.N + 1, 36)
    assert np.abs(out["X"][0][:, :27] - ref["X"][0]).max() < 1e-8 and np.abs(out["U"] - ref["U"]).max() < 1e-7
    assert np.abs(out["X"][0][:, 27:] - xo).max() < 1e-12


def test_moving_dynamic_obstacle_prediction_and_avoidance():
    """The obstacle states of the solution follow the constant-acceleration model from the observation, and the
    robot reacts to where the obstacle WILL be (rows linearised at the held guess + the known obstacle step)."""
    import oracle
    desc, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    slot = [i for i in range(desc.n_spheres) if desc.spheres[i].link == -1][0]
    p0 = np.array(list(desc.spheres[slot].offset))
    x0 = np.array(meta["x0"], dtype=float)
    target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (desc.N + 1, 1))
    r_ee = np.array(oracle.fk(desc, x0)["r"])
    far = p0 + 3.0 * (p0 - r_ee) / np.linalg.norm(p0 - r_ee)          # starts 3 m further away ...
    vel = (r_ee - far) / 2.0                                           # ... and reaches the tray in 2 s
    dyn, xo = _with_dynamic_obstacle(desc, slot, np.concatenate((far, vel, [0.0, 0.0, -0.1])))
    out = oracle.solve_batch(dyn, np.concatenate((x0, xo)), target)
    assert out["status"][0] in (0, 1) and out["stats"][0][3] == 1.0       # full step accepted
    t = desc.dt * np.arange(desc.N + 1)[:, None]
    pred = np.hstack((xo[:3] + t * xo[3:6] + 0.5 * t * t * xo[6:], xo[3:6] + t * xo[6:], np.tile(xo[6:], (desc.N + 1, 1))))
    assert np.abs(out["X"][0][:, 27:] - pred).max() < 1e-10
    parked, xs = _with_dynamic_obstacle(desc, slot, np.concatenate((far, np.zeros(6))))
    still = oracle.solve_batch(parked, np.concatenate((x0, xs)), target)
    assert np.abs(out["X"][0][:, :27] - still["X"][0][:, :27]).max() > 1e-3   # the approaching obstacle changes the plan


def _cubic_newton_literal(r, xo):
    """projectile_path_constraint.h:11-44, line by line."""
    r0, v0, g = xo[:3], xo[3:6], xo[6:]
    dr = r - r0
    a, b, c, d = g @ g, 3 * v0 @ g, 2 * (v0 @ v0 - dr @ g), -2 * dr @ v0
    x = 0.0
    for _ in range(10):
        f = a * x * x * x + b * x * x + c * x + d
        dfdx = 3 * a * x * x + 2 * b * x + c
        update = f / dfdx
        x = x - update
        if abs(update) < 1e-4:
            return x
    return x


def test_projectile_path_rows():
    """ProjectilePathConstraint (projectile_path_constraint.h:46-156): value against a brute-force closest approach,
    Jacobian (time of closest approach held fixed = envelope derivative) against finite differences over robot and
    obstacle states, the flag s, and the literal Newton iteration where it does NOT converge (ball thrown upwards)."""
    import oracle
    from _util import projectile_problem, projectile_throws
    d, meta, tray = projectile_problem()
    base, _ = problem_io.load_fixture("cfg4_thing_obstacles2")
    assert oracle.dims(d)["nx"] == 36 and oracle.dims(d)["n_ineq"] == oracle.dims(base)["n_ineq"] + 2
    x0 = np.array(meta["x0"], dtype=float)
    spheres = oracle.fk(d, np.concatenate((x0, np.zeros(9))))["spheres"]
    slots = [d.projectile_spheres[i] for i in range(2)]
    dist = [d.projectile_distances[i] for i in range(2)]
    tt = np.linspace(0, 3, 300001)
    for xo in projectile_throws(spheres[tray]):
        x = np.concatenate((x0, xo))
        pr = oracle.projectile(d, x)
        path = xo[None, :3] + tt[:, None] * xo[None, 3:6] + 0.5 * tt[:, None] ** 2 * xo[None, 6:]
        for i in range(2):
            dmin = np.linalg.norm(spheres[slots[i]][None] - path, axis=1).min()
            assert abs(pr["h"][i] - 0.2 / dist[i] * (dmin - dist[i])) < 1e-7
            assert abs(pr["tclose"][i] - max(0.0, _cubic_newton_literal(spheres[slots[i]], xo))) < 1e-12
        J = np.zeros_like(pr["J"])
        h = 1e-6
        for j in range(len(x)):
            xp, xm = x.copy(), x.copy()
            xp[j] += h
            xm[j] -= h
            J[:, j] = (oracle.projectile(d, xp)["h"] - oracle.projectile(d, xm)["h"]) / (2 * h)
        assert np.abs(pr["J"] - J).max() < 1e-6 and np.abs(pr["J"][:, 9:27]).max() == 0   # positions only
    # ball that has passed: closest approach clamps to t = 0 (":123 don't care about the past")
    xo = np.concatenate((spheres[tray] + [0.0, 1.0, 0.0], [0.0, 3.0, -1.0], [0.0, 0.0, -9.81]))
    pr = oracle.projectile(d, np.concatenate((x0, xo)))
    assert pr["tclose"][0] == 0.0 and abs(pr["h"][0] - 0.2 / 0.35 * (1.0 - 0.35)) < 1e-12
    # upward throw of the reference's own simulation (obstacles/dynamic.yaml:48-52): Newton leaves t = 0 towards a
    # far root and runs out of iterations — reproduced, not repaired
    xo = np.concatenate((spheres[tray] + [0.0, -2.0, 0.3], [0.0, 2.67, 3.68], [0.0, 0.0, -9.81]))
    pr = oracle.projectile(d, np.concatenate((x0, xo)))
    lit = _cubic_newton_literal(spheres[tray], xo)
    assert abs(pr["tclose"][0] - lit) < 1e-9 * max(1.0, abs(lit))
    # s = 0: rows vanish identically (value and Jacobian); s <= 0.5 never looks ahead
    d0, _, _ = projectile_problem(active=0.0)
    pr0 = oracle.projectile(d0, np.concatenate((x0, projectile_throws(spheres[tray])[0])))
    assert np.all(pr0["h"] == 0) and np.all(pr0["J"] == 0) and np.all(pr0["tclose"] == 0)


def test_projectile_constraint_shapes_the_plan():
    """With the flag raised the plan gives way to the ball's predicted path (soft rows: the violation shrinks), with
    the flag down it is the plan of the problem without the rows."""
    import oracle
    from _util import ballistic_prediction, projectile_problem, projectile_throws
    d, meta, tray = projectile_problem()
    d0, _, _ = projectile_problem(active=0.0)
    x0 = np.array(meta["x0"], dtype=float)
    spheres = oracle.fk(d, np.concatenate((x0, np.zeros(9))))["spheres"]
    target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (d.N + 1, 1))
    xo = projectile_throws(spheres[tray])[0]
    x = np.concatenate((x0, xo))
    Xw = np.hstack((np.tile(x0, (d.N + 1, 1)), ballistic_prediction(xo, d.N, d.dt)))[None]
    Uw = np.zeros((1, d.N, 13))
    on = oracle.solve_batch(d, x, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    off = oracle.solve_batch(d0, x, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    assert on["status"][0] == 0 and off["status"][0] == 0
    # obstacle states of an accepted full step = the ballistic prediction
    assert on["stats"][0, 3] == 1.0 and np.abs(on["X"][0][:, 27:] - Xw[0][:, 27:]).max() < 1e-9
    rows = lambda X: np.array([oracle.projectile(d, X[k])["h"] for k in range(1, d.N)])  # noqa: E731
    h_on, h_off = rows(on["X"][0]), rows(off["X"][0])
    assert h_off.min() < -0.05 and (np.minimum(h_on, 0) ** 2).sum() < 0.97 * (np.minimum(h_off, 0) ** 2).sum()
    # flag down = the problem without the rows (same robot trajectory as the plain dynamic-obstacle problem)
    import copy
    plain = copy.deepcopy(d0)
    plain.projectile_enabled = 0
    ref = oracle.solve_batch(plain, x, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    assert np.abs(ref["X"] - off["X"]).max() < 1e-6 and np.abs(ref["U"] - off["U"]).max() < 1e-5


def test_precision_study_modes():
    """oracle.qp_step_precision (tools/precision_lab.py): the fp64 mode is the oracle's interior-point step computed
    with the kernels' form of the Riccati recursion (partial Cholesky + Schur complement) instead of the dense gain
    recursion — same step; the fp32 mode stays within the kernels' stated tolerance; fp32 factors with an fp64
    iterate come closer."""
    import oracle
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    x0 = np.array(meta["x0"], dtype=float)
    target = np.tile(oracle.fk(desc, x0)["r"] + [0.2, 0.1, 0.1], (desc.N + 1, 1))
    X, U = np.tile(x0, (desc.N + 1, 1)), np.zeros((desc.N, desc.nu))
    dX, dU, info = oracle.qp_step(desc, target, X, U)
    r64, r32, rmx = (oracle.qp_step_precision(desc, target, X, U, m) for m in (0, 1, 2))
    assert r64["converged"] and r64["iters"] == info["iters"]
    assert np.abs(r64["dX"] - dX).max() < 1e-9 and np.abs(r64["dU"] - dU).max() < 1e-8
    rx = np.array(desc.state_ub[:27]) - np.array(desc.state_lb[:27])
    e32 = (np.abs(r32["dX"] - dX) / rx).max()
    emx = (np.abs(rmx["dX"] - dX) / rx).max()
    assert r32["converged"] and rmx["converged"] and r32["failed"] == 0
    assert e32 < 1e-2 and emx < 1e-4 and emx <= e32


def test_long_horizon_robust_planner_shape():
    """N = 100 knots with three SQP iterations (upright_robust/config/demos/_base.yaml:62-66): the Riccati recursion
    keeps the cost-to-go symmetric, so the 100-stage backward sweep stays positive definite (an unsymmetrised
    recursion breaks down near stage 60) and the longer horizon reproduces the 20-knot plan where they overlap in
    character: converged, dynamically consistent, objects balanced within the slack tolerance."""
    import copy
    import oracle
    from upright_b200 import workload
    base, meta = problem_io.load_fixture("cfg5_thing_robust8")
    desc = copy.deepcopy(base)
    desc.N, desc.sqp_iteration = 100, 3
    ee = lambda x: np.stack([oracle.fk(desc, xi)["r"] for xi in x])  # noqa: E731
    b = workload.sample_batch("cfg5_thing_robust8", desc, meta, 3, 11, ee)
    out = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    assert (out["status"] == 0).all() and (out["stats"][:, 7] == 3).all()
    X, U = out["X"], out["U"]
    assert X.shape == (3, 101, desc.nx) and np.all(np.isfinite(X)) and np.all(np.isfinite(U))
    A, Bm = _dynamics(desc)
    for i in range(3):
        gap = X[i, 1:] - X[i, :-1] @ A.T - U[i] @ Bm.T
        assert np.abs(gap).max() < 1e-6
        # three SQP iterations move the end of the plan towards the goal (soft terminal rows: not onto it)
        r_end, r_start = oracle.fk(desc, X[i, -1])["r"], oracle.fk(desc, X[i, 0])["r"]
        goal = b["target"][i, -1]
        assert np.linalg.norm(r_end - goal) < 0.8 * np.linalg.norm(r_start - goal)


def test_hopeless_hard_rows_end_the_qp_early():
    """cfg4 (hard object-dynamics rows): a few per cent of the seeded start states have inconsistent rows; their QPs
    used to run to the 30-iteration cap and now end once the linear rate of the multiplier method shows that the
    tolerance is out of reach — still reported as not converged — while every other
    instance converges exactly as before (the criterion never fires on a row residual that shrinks tenfold per
    iteration)."""
    import oracle
    from upright_b200 import workload
    name = "cfg4_thing_obstacles2"
    desc, meta = problem_io.load_fixture(name)
    ee = lambda x: np.stack([oracle.fk(desc, xi)["r"] for xi in x])  # noqa: E731
    mg = lambda x: np.array([oracle.linearize(desc, xi, np.zeros(desc.nu))["hobs"] for xi in x])  # noqa: E731
    b = workload.sample_batch(name, desc, meta, 128, 1234, ee, margin_fn=mg)
    out = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    it = out["stats"][:, 0]
    capped = out["status"] == 1
    assert 1 <= capped.sum() <= 8 and (out["status"][~capped] == 0).all()
    assert it[capped].max() <= 12 and it[capped].min() >= 6          # ended by the rate test, not by the cap of 30
    assert it[~capped].max() <= 10
    # the remaining equality violation of those instances is far from the tolerance: they were never going to converge
    assert out["stats"][capped, 5].min() > 1e-2
