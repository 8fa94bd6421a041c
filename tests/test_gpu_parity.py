"""Parity of the CUDA path (through the C ABI) against the fp64 CPU oracle on
the same seeded inputs, plus size-independent properties at the full BASELINE
batch sizes.

Stated tolerances:
  fp64 kernels    : |X - X_oracle|, |U - U_oracle| <= 1e-7 (same algorithm, same arithmetic width)
  product kernels : ("f32": Riccati recursion in fp32; iterate, residuals, linearisation and force block in fp64)
                    errors scaled by the limit range of each component
                    (state_ub - state_lb, input_ub - input_lb, force_ub - force_lb):
                    worst instance <= 1e-3, median <= 1e-4, same interior-point iteration counts;
                    linearisation Jacobians (stored in float) abs 1e-6.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _util import has_cuda  # noqa: E402

pytestmark = pytest.mark.gpu
if not has_cuda():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import torch  # noqa: E402

import oracle  # noqa: E402
from upright_b200 import problem_io, workload  # noqa: E402
from upright_b200.engine import BatchedMPC  # noqa: E402

CFGS = list(problem_io.FIXTURES)
_cache = {}


def engine(name, prec):
    key = (name, prec)
    if key not in _cache:
        desc, meta = problem_io.load_fixture(name)
        _cache[key] = (BatchedMPC(desc, prec), desc, meta)
    return _cache[key]


def batch_for(name, B, seed):
    mpc, desc, meta = engine(name, "f64")
    ee = lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu)))  # noqa: E731
    mg = (lambda x: mpc.eval("obstacle_avoidance", x, np.zeros((x.shape[0], mpc.nu)))) if desc.obstacles_enabled else None
    return workload.sample_batch(name, desc, meta, B, seed, ee, margin_fn=mg)


def ranges(desc):
    nq, nx, nu = desc.nq, desc.nx, desc.nu
    rx = np.array(desc.state_ub[:nx]) - np.array(desc.state_lb[:nx])
    ru = np.concatenate([np.array(desc.input_ub[:nq]) - np.array(desc.input_lb[:nq]),
                         np.full(nu - nq, desc.force_ub - desc.force_lb)])
    return rx, ru


@pytest.mark.parametrize("name", CFGS)
def test_probes_match_oracle(name):
    mpc, desc, meta = engine(name, "f64")
    rng = np.random.default_rng(0)
    M = 16
    x = np.tile(meta["x0"], (M, 1)) + 0.2 * rng.standard_normal((M, desc.nx))
    u = rng.standard_normal((M, desc.nu))
    r = mpc.eval("end_effector_position", x, u)
    g = mpc.eval("object_dynamics", x, u)
    for m in range(M):
        lin = oracle.linearize(desc, x[m], u[m])
        assert np.allclose(r[m], lin["r"], atol=1e-12)
        assert np.allclose(g[m], lin["g"], atol=1e-11)
    if desc.n_fric:
        h = mpc.eval("contact_forces", x, u)
        for m in range(M):
            assert np.allclose(h[m], oracle.linearize(desc, x[m], u[m])["hfric"], atol=1e-12)
    if desc.n_obs:
        h = mpc.eval("obstacle_avoidance", x, u)
        for m in range(M):
            assert np.allclose(h[m], oracle.linearize(desc, x[m], u[m])["hobs"], atol=1e-12)
    # Jacobian probes (what the reference's CppAD tapes return): object dynamics [dg/dx | dg/du], tool position dr/dq
    J = mpc.eval("end_effector_jacobian", x, u).reshape(M, 3, desc.nq)
    for m in range(M):
        assert np.allclose(J[m], oracle.linearize(desc, x[m], u[m])["Jp"], atol=1e-12)
    if mpc.n_eq:
        G = mpc.eval("object_dynamics_jacobian", x, u).reshape(M, mpc.n_eq, desc.nx + desc.nu)
        for m in range(0, M, 5):
            lin = oracle.linearize(desc, x[m], u[m])
            assert np.allclose(G[m][:, :desc.nx], lin["C"], atol=1e-10)
            assert np.allclose(G[m][:, desc.nx + desc.nq:], lin["Df"], atol=1e-12) and np.all(G[m][:, desc.nx:desc.nx + desc.nq] == 0)


@pytest.mark.parametrize("name", CFGS)
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_linearisation_blocks(name, prec):
    mpc, desc, meta = engine(name, prec)
    b = batch_for(name, 4, 3)
    dt = mpc.torch_dtype
    mpc.set_option("stop_after", 1)
    try:
        mpc.solve_device(torch.tensor(b["x0"], dtype=dt, device="cuda"), torch.tensor(b["target"], dtype=dt, device="cuda"),
                         None if b["body_params"] is None else torch.tensor(b["body_params"], dtype=dt, device="cuda"))
        torch.cuda.synchronize()
        N, nx, nq, neq, nfc = desc.N, desc.nx, desc.nq, mpc.n_eq, desc.nu - desc.nq
        blk = {n: mpc.workspace_block(4, n, c) for n, c in (("LJP", (N + 1) * 3 * nq), ("LR", (N + 1) * 3), ("LC", N * neq * nx),
                                                            ("LG", N * neq), ("DF", neq * nfc), ("LHO", (N + 1) * desc.n_obs),
                                                            ("LJO", (N + 1) * desc.n_obs * nq))}
    finally:
        mpc.set_option("stop_after", 0)
    # the product kernels linearise in double and store the Jacobians in float (the values the residuals are built
    # from — LG, LR — stay double): 1e-6 absolute covers the storage rounding of entries up to ~10
    tol = 1e-12 if prec == "f64" else 1e-6
    vtol = 1e-12 if prec == "f64" else 2e-6      # fp32 kernels: x0 itself is rounded to float on the way in
    for inst in range(4):
        bp = None if b["body_params"] is None else b["body_params"][inst]
        lin = oracle.linearize(desc, b["x0"][inst], np.zeros(desc.nu), bp)
        k = 5  # cold start: every knot is linearised at x0, u = 0
        Jp = blk["LJP"][inst, k * 3 * nq: (k + 1) * 3 * nq].reshape(3, nq)
        assert np.allclose(Jp, lin["Jp"], atol=tol)
        assert np.allclose(blk["LR"][inst, 3 * k: 3 * k + 3], lin["r"], atol=vtol)
        if neq:
            rows = blk["LC"][inst, k * neq * nx: (k + 1) * neq * nx].reshape(neq, nx)
            assert np.allclose(rows, lin["C"], atol=10 * tol)
            assert np.allclose(blk["LG"][inst, k * neq: (k + 1) * neq], lin["g"], atol=10 * vtol)
            Df = blk["DF"][inst].reshape(neq, nfc)
            assert np.allclose(Df, lin["Df"], atol=10 * tol)
        if desc.n_obs:
            assert np.allclose(blk["LHO"][inst, k * desc.n_obs: (k + 1) * desc.n_obs], lin["hobs"], atol=vtol)
            Jo = blk["LJO"][inst, k * desc.n_obs * nq: (k + 1) * desc.n_obs * nq].reshape(-1, nq)
            assert np.allclose(Jo, lin["Jobs"], atol=tol)


@pytest.mark.parametrize("name", CFGS)
def test_full_solve_fp64_matches_oracle(name):
    mpc, desc, meta = engine(name, "f64")
    b = batch_for(name, 32, 11)
    out = mpc.solve(b["x0"], b["target"], b["body_params"], want_gains=True)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"], want_gains=True)
    assert np.array_equal(out["status"], ref["status"])
    good = ref["status"] == 0
    assert good.mean() >= 0.75, f"only {good.mean():.2f} of the oracle solves converged"
    assert np.array_equal(out["stats"][good, 0], ref["stats"][good, 0])          # same IPM iteration counts
    assert np.abs(out["X"][good] - ref["X"][good]).max() < 1e-7
    # inputs: 1e-6 — the kernels eliminate the forces stage by stage, the oracle factorises the dense stage matrices;
    # with friction pyramids (cfg3) the weakly determined tangential forces differ by 1.1e-7 at equal iteration counts
    assert np.abs(out["U"][good] - ref["U"][good]).max() < 1e-6
    assert np.allclose(out["stats"][good, 1:4], ref["stats"][good, 1:4], rtol=1e-7, atol=1e-9)
    # feedback gains: hard-equality configurations carry proximal weights up to rho_hard/|a|^2 ~ 1e10 in the
    # stage Hessians, so two fp64 summation orders agree to ~1e-4 relative there
    ktol = 1e-3 if not desc.slacks.enabled else 1e-6
    assert np.abs(out["K"][good] - ref["K"][good]).max() < ktol * max(1.0, np.abs(ref["K"][good]).max())


@pytest.mark.parametrize("name", CFGS)
def test_full_solve_fp32_within_stated_tolerance(name):
    mpc, desc, meta = engine(name, "f32")
    B = 128 if desc.nu <= 13 else 48
    b = batch_for(name, B, 21)
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    assert (out["status"] != 3).all()                      # no breakdown of the fp32 factorisation
    assert (out["status"] == ref["status"]).mean() >= 0.97
    good = (ref["status"] == 0) & (out["status"] == 0)
    assert good.mean() >= 0.7
    # Same interior-point iteration counts on (almost) every instance.  The convergence test is a threshold
    # (mu <= 2 mu_target): an instance that ends an iteration within rounding of it (cfg5, seed 21, instance 36:
    # mu = 2.0001e-7 in fp64, 1.9992e-7 in the product kernels — the CPU study of the same arithmetic,
    # tools/precision_lab2.py, reproduces it) stops one iteration apart and returns the central-path point of that
    # iteration, ~1e-3 of the range away in the weakly determined force directions.  Those instances are held to 1e-2.
    same = out["stats"][:, 0] == ref["stats"][:, 0]
    assert same[good].mean() >= 0.97
    rx, ru = ranges(desc)
    ex = (np.abs(out["X"] - ref["X"]) / rx).reshape(B, -1).max(1)
    eu = (np.abs(out["U"] - ref["U"]) / ru).reshape(B, -1).max(1)
    e_all = np.maximum(ex, eu)
    e = e_all[good & same]
    print(f"{name}: fp32 scaled error max {e.max():.2e} p95 {np.percentile(e, 95):.2e} median {np.median(e):.2e}"
          f" ({int((good & ~same).sum())} instance(s) one iteration apart, max {e_all[good & ~same].max() if (good & ~same).any() else 0:.2e})")
    assert e.max() <= 1e-3 and np.median(e) <= 1e-4
    assert (good & ~same).sum() == 0 or e_all[good & ~same].max() <= 1e-2
    good = good & same
    # constraint residuals of the reference's own functions on the returned trajectory
    assert np.allclose(out["stats"][good, 2], ref["stats"][good, 2], rtol=2e-3, atol=2e-4)   # violation
    assert np.allclose(out["stats"][good, 1], ref["stats"][good, 1], rtol=1e-3, atol=1e-4)   # cost


def test_device_path_equals_host_path():
    mpc, desc, meta = engine("cfg2_thing_demo", "f32")
    b = batch_for("cfg2_thing_demo", 64, 5)
    host = mpc.solve(b["x0"], b["target"], b["body_params"])
    dev = mpc.solve_device(torch.tensor(b["x0"], dtype=torch.float32, device="cuda"),
                           torch.tensor(b["target"], dtype=torch.float32, device="cuda"),
                           torch.tensor(b["body_params"], dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(dev["X"].cpu().numpy().astype(np.float64), host["X"])
    assert np.array_equal(dev["U"].cpu().numpy().astype(np.float64), host["U"])
    assert np.array_equal(dev["status"].cpu().numpy(), host["status"])


def test_full_batch_properties_cfg2():
    """BASELINE size (4096 instances): properties that need no oracle solve of that size."""
    name = "cfg2_thing_demo"
    mpc, desc, meta = engine(name, "f32")
    B = workload.BASELINE_BATCH[name]
    b = batch_for(name, B, 99)
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    again = mpc.solve(b["x0"], b["target"], b["body_params"])
    assert np.array_equal(out["X"], again["X"]) and np.array_equal(out["U"], again["U"])    # deterministic
    assert set(np.unique(out["status"])) <= {0, 1, 2}
    assert (out["status"] == 0).mean() > 0.98
    assert np.all(np.isfinite(out["X"])) and np.all(np.isfinite(out["U"]))
    X, U = out["X"], out["U"]
    assert np.allclose(X[:, 0], b["x0"], atol=1e-6)                                          # x_0 is the observation
    # accepted full steps are dynamically consistent: x_{k+1} = A x_k + B u_k (exact discretisation)
    nq, dt = desc.nq, desc.dt
    full = out["stats"][:, 3] == 1.0
    q, v, a, j = X[:, :-1, :nq], X[:, :-1, nq:2 * nq], X[:, :-1, 2 * nq:], U[:, :, :nq]
    gq = q + dt * v + 0.5 * dt * dt * a + dt**3 / 6 * j - X[:, 1:, :nq]
    gv = v + dt * a + 0.5 * dt * dt * j - X[:, 1:, nq:2 * nq]
    ga = a + dt * j - X[:, 1:, 2 * nq:]
    gap = np.maximum(np.abs(gq).max((1, 2)), np.maximum(np.abs(gv).max((1, 2)), np.abs(ga).max((1, 2))))
    assert full.mean() > 0.9 and gap[full].max() < 2e-4
    # reported cost / violation are the oracle's performance index of the returned trajectory
    idx = np.random.default_rng(0).choice(B, 64, replace=False)
    for i in idx:
        pf = oracle.performance(desc, b["target"][i], X[i], U[i], b["body_params"][i])
        assert np.isclose(pf["cost"], out["stats"][i, 1], rtol=2e-4, atol=1e-5)
        assert np.isclose(pf["violation"], out["stats"][i, 2], rtol=2e-3, atol=1e-4)
    # a second (warm-started) SQP iteration lowers the constraint violation on almost all instances
    warm = mpc.solve(b["x0"], b["target"], b["body_params"], X=X, U=U, warm=True)
    assert (warm["stats"][:, 2] < out["stats"][:, 2]).mean() > 0.9
    # sampled instances agree with the oracle
    sub = idx[:16]
    ref = oracle.solve_batch(desc, b["x0"][sub], b["target"][sub], b["body_params"][sub])
    rx, ru = ranges(desc)
    assert (np.abs(X[sub] - ref["X"]) / rx).max() < 1e-2 and (np.abs(U[sub] - ref["U"]) / ru).max() < 1e-2


@pytest.mark.parametrize("name", ["cfg3_thing_box_arch", "cfg5_thing_robust8"])
def test_full_batch_runs(name):
    mpc, desc, meta = engine(name, "f32")
    B = 1024
    b = batch_for(name, B, 17)
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    assert np.all(np.isfinite(out["X"])) and (out["status"] <= 1).mean() > 0.97
    assert (out["status"] == 0).mean() > 0.9


def test_controller_manager_drop_in():
    """mpc_sim.py-style closed loop through the reference-facing surface
    (manager.py:156-176 semantics) with the double-integrator plant of
    mpc_sim.py:148-155."""
    from upright_b200.manager import ControllerManager
    from upright_b200.trajectory import DoubleIntegrator
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    cfg = meta["controller_config"]
    mgr = ControllerManager.from_config(cfg, x0=np.array(meta["x0"]))
    dims = mgr.model.settings.dims
    nq = dims.robot.q
    assert mgr.mpc.getStateDim() == 27 and mgr.mpc.getInputDim() == 13
    x = np.array(meta["x0"], dtype=float)
    integ = DoubleIntegrator(nq)
    dt_sim, t = 0.01, 0.0
    r0 = mgr.mpc._engine.eval("end_effector_position", x, np.zeros(13))[0]
    goal = r0 + np.array([-0.25, 0.5, 0.25])  # thing_demo.yaml:55
    dist0 = np.linalg.norm(goal - r0)
    for step in range(150):
        xd, u = mgr.step(t, x)
        assert np.all(np.isfinite(u))
        q, v, a = x[:nq], x[nq:2 * nq], x[2 * nq:]
        v_new, a_new = integ.integrate(v, a, u[:nq], dt_sim)
        q = q + dt_sim * v + 0.5 * dt_sim**2 * a + dt_sim**3 / 6 * u[:nq]
        x = np.concatenate((q, v_new, a_new))
        t += dt_sim
    assert len(mgr.replanning_times) >= 100                     # replans every min_policy_update_time
    r = mgr.mpc._engine.eval("end_effector_position", x, np.zeros(13))[0]
    assert np.linalg.norm(goal - r) < 0.6 * dist0               # moving toward the waypoint
    g = mgr.mpc.getStateInputEqualityConstraintValue("object_dynamics", t, x, u)
    assert np.abs(g).max() < 0.5                                # balancing residual stays small
    ts, xs, us = mgr.get_mpc_trajectory()
    assert ts.shape == (21,) and xs.shape == (21, 27) and us.shape == (21, 13)


def _host_closed_loop(settings, targets, x0, n_steps, sim_dt, body_params=None):
    """numpy mirror of the device rollout: BatchedControllerManager.step (replan gate, warm-start shift,
    policy evaluation in manager._RecedingHorizon) + the tracking law and exact triple-integrator plant."""
    from upright_b200.manager import BatchedControllerManager
    mgr = BatchedControllerManager(settings, targets, body_params=body_params)
    nq = settings.dims.robot.q
    tr = settings.tracking
    x = np.array(x0, dtype=float)
    xs, us = [], []
    for s in range(n_steps):
        t = sim_dt * s
        xd, u = mgr.step(t, x)
        q, v, a = x[:, :nq], x[:, nq:2 * nq], x[:, 2 * nq:]
        e = xd - x
        ucmd = tr.kp * e[:, :nq] + tr.kv * e[:, nq:2 * nq] + tr.ka * e[:, 2 * nq:] + u[:, :nq]
        xs.append(x.copy())
        us.append(ucmd.copy())
        h = sim_dt
        x = np.hstack((q + h * v + 0.5 * h * h * a + h**3 / 6 * ucmd, v + h * a + 0.5 * h * h * ucmd, a + h * ucmd))
    return np.stack(xs, 1), np.stack(us, 1), x, len(mgr.replanning_times)


@pytest.mark.parametrize("feedback", [True, False])
def test_device_closed_loop_matches_host_manager(feedback):
    """ub_closed_loop (targets, warm-start shift, policy evaluation, plant on the device) against the
    host-side numpy mirror of manager.py:156-176 / mpc_sim.py:118-160 driving the same solve kernel."""
    from upright_b200.manager import BatchedControllerManager
    from upright_b200.settings import ControllerSettings, TargetTrajectories
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    cfg = meta["controller_config"]
    settings = ControllerSettings(config=cfg, x0=np.array(meta["x0"]))
    settings.sqp.use_feedback_policy = feedback
    B, sim_dt, n_steps = 5, 0.005, 50
    rng = np.random.default_rng(3)
    x0 = np.tile(np.array(meta["x0"], dtype=float), (B, 1))
    x0[:, :9] += rng.uniform(-0.1, 0.1, (B, 9))
    probe = BatchedControllerManager(settings, [None] * B)
    r0 = probe.engine.eval("end_effector_position", x0, np.zeros((B, 13)))
    quat = np.array([0.0, 0.0, 0.0, 1.0])
    targets = []
    for b in range(B):
        wp = [np.r_[r0[b] + d, quat, 0.0] for d in ([0.0, 0.0, 0.0], [-0.25, 0.5, 0.25], [0.2, -0.3, 0.1])]
        targets.append(TargetTrajectories([0.0, 0.12, 0.5], wp, [np.zeros(13)] * 3))
    xs_h, us_h, xf_h, nrep_h = _host_closed_loop(settings, targets, x0, n_steps, sim_dt)
    mgr = BatchedControllerManager(settings, targets)
    out = mgr.rollout(x0, n_steps * sim_dt, sim_dt)
    assert out["n_replans"] == nrep_h and 20 <= nrep_h <= 25   # same double comparisons on both sides
    assert out["status_counts"].sum() == nrep_h * B and (out["status_counts"][:, 3] == 0).all()
    assert np.isfinite(out["xs"]).all() and np.isfinite(out["us"]).all()
    # same fp32 solver on both sides; the only differences are float vs double bookkeeping between replans
    scale = np.abs(us_h).max()
    assert np.abs(out["us"] - us_h).max() <= 2e-2 * scale, np.abs(out["us"] - us_h).max() / scale
    assert np.abs(out["xs"] - xs_h).max() <= 1e-3
    assert np.abs(out["x_final"] - xf_h).max() <= 1e-3
    # the robot actually moves toward the waypoint
    assert np.abs(out["x_final"][:, :9] - x0[:, :9]).max() > 1e-3


def test_device_closed_loop_log_stride_and_errors():
    from upright_b200.engine import BatchedMPC
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    mpc = BatchedMPC(desc, "f32")
    x0 = np.tile(np.array(meta["x0"], dtype=float), (3, 1))
    r0 = mpc.eval("end_effector_position", x0, np.zeros((3, 13)))
    goal = (r0 + np.array([-0.25, 0.5, 0.25]))[:, None, :]
    full = mpc.closed_loop(x0, [0.0], goal, 40, 0.005, 0.01)
    thin = mpc.closed_loop(x0, [0.0], goal, 40, 0.005, 0.01, log_stride=4)
    assert thin["xs"].shape == (3, 10, 27) and thin["us"].shape == (3, 10, 9)
    assert np.array_equal(thin["xs"], full["xs"][:, ::4]) and np.array_equal(thin["x_final"], full["x_final"])
    nolog = mpc.closed_loop(x0, [0.0], goal, 40, 0.005, 0.01, log=False)
    assert nolog["xs"] is None and np.array_equal(nolog["x_final"], full["x_final"])
    with pytest.raises(RuntimeError):
        mpc.closed_loop(x0, [0.0, 0.0], np.repeat(goal, 2, axis=1), 10, 0.005, 0.01)   # times must increase
    with pytest.raises(RuntimeError):
        mpc.closed_loop(x0, [0.0], goal, 0, 0.005, 0.01)


def test_no_fp32_breakdowns_and_rescue_flag_is_inert():
    """Round 1's fp32 kernels lost positive definiteness on ~0.3 % of the cfg3 instances (status NAN) and needed the
    fp64 re-solve of UB_RESCUE_F64.  The reduced-stage kernels factorise no penalty-weighted matrix: no instance
    breaks down, and the flag (kept in the ABI) changes nothing."""
    name = "cfg3_thing_box_arch"
    mpc, desc, meta = engine(name, "f32")
    b = batch_for(name, 2048, 1234)
    plain = mpc.solve(b["x0"], b["target"], b["body_params"])
    assert (plain["status"] != 3).all() and np.isfinite(plain["X"]).all()
    resc = mpc.solve(b["x0"], b["target"], b["body_params"], rescue=True)
    assert np.array_equal(resc["X"], plain["X"]) and np.array_equal(resc["status"], plain["status"])


def test_light_object_with_soft_rows_solves_in_the_product_kernels():
    """The documented fp32 conditioning limit of round 1: the projectile problem with cfg4's
    20 g object and soft object-dynamics rows, warm-started.  The penalty Z / (6 m^2) ~ 4e4 against force weights of
    1e-4 never reaches a factorisation now (only its reciprocal does): same status and iteration counts as the oracle,
    result within the stated tolerance."""
    from _util import ballistic_prediction, projectile_problem, projectile_throws
    d, meta, tray = projectile_problem(light_object=True)
    x0r = np.array(meta["x0"], dtype=float)
    throws = projectile_throws(oracle.fk(d, np.concatenate((x0r, np.zeros(9))))["spheres"][tray])
    Bn = len(throws)
    x0 = np.hstack((np.tile(x0r, (Bn, 1)), throws))
    target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (Bn, d.N + 1, 1))
    Xw = np.stack([np.hstack((np.tile(x0r, (d.N + 1, 1)), ballistic_prediction(xo, d.N, d.dt))) for xo in throws])
    Uw = np.zeros((Bn, d.N, 13))
    mpc = BatchedMPC(d, "f32")
    out = mpc.solve(x0, target, None, X=Xw.copy(), U=Uw.copy(), warm=True)
    ref = oracle.solve_batch(d, x0, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    cold = oracle.solve_batch(d, x0, target)
    assert (out["status"] == ref["status"]).all() and (out["status"] != 3).all()
    rx, ru = ranges(d)
    ex = (np.abs(out["X"][:, :, :27] - ref["X"][:, :, :27]) / rx).max()
    eu = (np.abs(out["U"] - ref["U"]) / ru).max()
    print(f"light object: scaled error X {ex:.2e} U {eu:.2e}")
    assert ex <= 1e-3 and eu <= 1e-3
    assert np.abs(cold["X"] - ref["X"]).max() > 1e-2          # the warm start matters


def test_end_effector_box_constraint_parity():
    """EndEffectorBoxConstraint rows (end_effector_box_constraint.h:46-76): probe values, fp64 kernels against
    the oracle to 1e-7, fp32 within the stated tolerance, and the rows are active."""
    import copy
    base, meta = problem_io.load_fixture("cfg2_thing_demo")
    desc = copy.deepcopy(base)
    desc.ee_box_enabled = 1
    desc.ee_box_lower[:] = [-5.0, -5.0, -0.02]
    desc.ee_box_upper[:] = [5.0, 5.0, 0.02]
    b = batch_for("cfg2_thing_demo", 12, 77)
    m64, m32, mfree = BatchedMPC(desc, "f64"), BatchedMPC(desc, "f32"), BatchedMPC(base, "f64")
    assert m64.n_ineq == mfree.n_ineq + 6
    # probe
    tg0 = b["target"][:, 0, :]
    r0 = m64.eval("end_effector_position", b["x0"], np.zeros((12, 13)))
    h = m64.eval("end_effector_box_constraint", b["x0"], np.zeros((12, 13)), target=tg0)
    assert h.shape == (12, 6)
    assert np.allclose(h[:, :3], tg0 + np.array([5.0, 5.0, 0.02]) - r0, atol=1e-12)
    assert np.allclose(h[:, 3:], r0 - tg0 - np.array([-5.0, -5.0, -0.02]), atol=1e-12)
    assert mfree.eval("end_effector_box_constraint", b["x0"], np.zeros((12, 13)), target=tg0).shape == (12, 0)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    o64 = m64.solve(b["x0"], b["target"], b["body_params"])
    assert (o64["status"] == ref["status"]).all()
    assert np.abs(o64["X"] - ref["X"]).max() < 1e-7 and np.abs(o64["U"] - ref["U"]).max() < 1e-7
    free = mfree.solve(b["x0"], b["target"], b["body_params"])
    moved = np.abs(o64["X"] - free["X"]).reshape(12, -1).max(axis=1)
    assert (moved > 1e-3).sum() >= 6          # goals with |dz| > 0.02 make the rows active
    o32 = m32.solve(b["x0"], b["target"], b["body_params"])
    rx, ru = ranges(desc)
    ex = (np.abs(o32["X"] - ref["X"]) / rx).reshape(12, -1).max(axis=1)
    eu = (np.abs(o32["U"] - ref["U"]) / ru).reshape(12, -1).max(axis=1)
    ok = o32["status"] == ref["status"]
    assert ok.mean() >= 0.9 and np.median(ex[ok]) <= 3e-4 and ex[ok].max() <= 1e-2 and eu[ok].max() <= 1e-2


def test_inertial_alignment_cost_parity():
    """InertialAlignmentCostGaussNewton (inertial_alignment.cpp:90-163): probe value, linearisation blocks and the
    full solve of the fp64 kernels against the oracle (1e-7), fp32 within the stated tolerance."""
    import copy
    from upright_b200 import geometry as geo
    base, meta = problem_io.load_fixture("cfg2_thing_demo")
    desc = copy.deepcopy(base)
    desc.ia_cost_enabled = 1
    desc.ia_cost_weight = 50.0
    desc.ia_span[:] = geo.plane_span([0, 0, 1]).reshape(6)
    b = batch_for("cfg2_thing_demo", 10, 91)
    m64, m32 = BatchedMPC(desc, "f64"), BatchedMPC(desc, "f32")
    # probe: 1/2 w e'e at a moving state
    rng = np.random.default_rng(2)
    x = b["x0"].copy()
    x[:, 9:] = 0.3 * rng.standard_normal((10, 18))
    S = np.array(list(desc.ia_span)).reshape(2, 3)
    val = m64.eval("inertial_alignment_cost", x, np.zeros((10, 13)))[:, 0]
    for i in range(10):
        k = oracle.fk(desc, x[i])
        e = S @ (np.array(k["C"]).reshape(3, 3).T @ (np.array(k["a"]) - np.array(list(desc.gravity)))) / 9.81
        assert val[i] == pytest.approx(0.5 * 50.0 * e @ e, rel=1e-10, abs=1e-14)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    o64 = m64.solve(b["x0"], b["target"], b["body_params"])
    assert (o64["status"] == ref["status"]).all()
    assert np.abs(o64["X"] - ref["X"]).max() < 1e-7 and np.abs(o64["U"] - ref["U"]).max() < 1e-7
    assert np.abs(o64["stats"][:, 1] - ref["stats"][:, 1]).max() < 1e-8           # cost incl. the alignment term
    plain = BatchedMPC(base, "f64").solve(b["x0"], b["target"], b["body_params"])
    assert np.abs(o64["X"] - plain["X"]).max() > 1e-4                               # the cost acts
    o32 = m32.solve(b["x0"], b["target"], b["body_params"])
    rx, ru = ranges(desc)
    ex = (np.abs(o32["X"] - ref["X"]) / rx).reshape(10, -1).max(axis=1)
    eu = (np.abs(o32["U"] - ref["U"]) / ru).reshape(10, -1).max(axis=1)
    ok = o32["status"] == ref["status"]
    assert ok.mean() >= 0.9 and np.median(ex[ok]) <= 3e-4 and ex[ok].max() <= 1e-2 and eu[ok].max() <= 1e-2


@pytest.mark.parametrize("mode", ["plain", "angular", "fixed"])
def test_inertial_alignment_constraint_parity(mode):
    """InertialAlignmentConstraint rows (inertial_alignment.cpp:7-53), three variants: probe values against the
    literal numpy formula, full solve of the fp64 kernels against the oracle (1e-7), fp32 within tolerance."""
    import copy
    from upright_b200 import geometry as geo
    base, meta = problem_io.load_fixture("cfg2_thing_demo")
    desc = copy.deepcopy(base)
    desc.ia_constraint_enabled = 1
    desc.ia_alpha = 0.05
    desc.ia_normal[:] = [0.0, 0.0, 1.0]
    desc.ia_span[:] = geo.plane_span([0, 0, 1]).reshape(6)
    desc.ia_com[:] = [0.02, -0.01, 0.15]
    desc.ia_use_angular_acceleration = int(mode == "angular")
    desc.ia_align_with_fixed_vector = int(mode == "fixed")
    b = batch_for("cfg2_thing_demo", 8, 17)
    m64, m32 = BatchedMPC(desc, "f64"), BatchedMPC(desc, "f32")
    assert m64.n_ineq == BatchedMPC(base, "f64").n_ineq + 5
    rng = np.random.default_rng(4)
    x = b["x0"].copy()
    x[:, 9:] = 0.3 * rng.standard_normal((8, 18))
    hv = m64.eval("inertial_alignment_constraint", x, np.zeros((8, 13)))
    S, n, com = np.array(list(desc.ia_span)).reshape(2, 3), np.array([0, 0, 1.0]), np.array(list(desc.ia_com))
    sk = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])  # noqa: E731
    for i in range(8):
        k = oracle.fk(desc, x[i])
        C, w, al = np.array(k["C"]).reshape(3, 3), np.array(k["w"]), np.array(k["alpha"])
        a = C.T @ (np.array(k["a"]) - np.array(list(desc.gravity)))
        if mode == "angular":
            a = a + (sk(al) + sk(w) @ sk(w)) @ C @ com
        elif mode == "fixed":
            a = C.T @ n
        an, at = n @ a, S @ a
        want = [an, 0.05 * an - at[0] - at[1], 0.05 * an - at[0] + at[1], 0.05 * an + at[0] - at[1], 0.05 * an + at[0] + at[1]]
        assert np.allclose(hv[i], want, atol=1e-10)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    o64 = m64.solve(b["x0"], b["target"], b["body_params"])
    assert (o64["status"] == ref["status"]).all() and (o64["stats"][:, 0] == ref["stats"][:, 0]).all()
    # the linearisation agrees to 1e-15 (tools/ia_probe.py); in the "fixed" variant the start states violate the
    # tilt rows, the QP needs 11-15 interior-point iterations and the structured (kernel) vs dense (oracle)
    # algebra differ by 4e-6 at the same iteration counts
    tol = 2e-5 if mode == "fixed" else 1e-7
    assert np.abs(o64["X"] - ref["X"]).max() < tol and np.abs(o64["U"] - ref["U"]).max() < 100 * tol
    o32 = m32.solve(b["x0"], b["target"], b["body_params"])
    rx, ru = ranges(desc)
    ok = o32["status"] == ref["status"]
    ex = (np.abs(o32["X"] - ref["X"]) / rx).reshape(8, -1).max(axis=1)
    assert ok.mean() >= 0.75 and ex[ok].max() <= 1e-2


def _dyn_obstacle_desc(moving):
    """cfg4 with its first world sphere riding on a dynamic obstacle (+9 states)."""
    import copy
    desc, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    d = copy.deepcopy(desc)
    slot = [i for i in range(d.n_spheres) if d.spheres[i].link == -1][0]
    d.spheres[slot].link = -2
    d.n_dynamic_obstacles = 1
    p0 = np.array(list(d.spheres[slot].offset))
    return desc, d, meta, slot, p0


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_dynamic_obstacle_kernels_match_full_state_oracle(prec):
    """Dynamic obstacles (obstacle_constraint.h:8-43): the kernels eliminate the uncontrolled obstacle states
    analytically, the oracle carries them as genuine QP states (as the reference does) — same trajectories."""
    desc, dyn, meta, slot, p0 = _dyn_obstacle_desc(True)
    b = batch_for("cfg4_thing_obstacles2", 8, 5)
    mpc = BatchedMPC(dyn, prec)
    assert mpc.nx == 36 and mpc.nx_robot == 27
    r_ee = mpc.eval("end_effector_position", np.hstack((b["x0"], np.zeros((8, 9)))), np.zeros((8, mpc.nu)))
    # the obstacle starts 2 m behind the static sphere's place and heads for a point 0.7 m beside the tray
    dirn = (p0 - r_ee) / np.linalg.norm(p0 - r_ee, axis=1, keepdims=True)
    lateral = np.cross(dirn, [0.0, 0.0, 1.0])
    lateral /= np.linalg.norm(lateral, axis=1, keepdims=True)
    far = p0 + 2.0 * dirn
    vel = (r_ee + 0.7 * lateral - far) / 3.0
    xo = np.hstack((far, vel, np.zeros((8, 3))))
    xo[0] = np.concatenate((p0, np.zeros(6)))                    # instance 0: obstacle at rest where the sphere was
    x0 = np.hstack((b["x0"], xo))
    # probe: distances use the obstacle position carried in x
    h = mpc.eval("obstacle_avoidance", x0, np.zeros((8, mpc.nu)))
    href = np.array([oracle.linearize(dyn, x0[i], np.zeros(mpc.nu))["hobs"] for i in range(8)])
    assert np.allclose(h, href, atol=1e-10)
    ref = oracle.solve_batch(dyn, x0, b["target"], b["body_params"])
    out = mpc.solve(x0, b["target"], b["body_params"])
    ok = ref["status"] == 0
    assert ok.sum() >= 5
    if prec == "f64":
        assert (out["status"] == ref["status"]).all()
        assert np.abs(out["X"][ok] - ref["X"][ok]).max() < 1e-7 and np.abs(out["U"][ok] - ref["U"][ok]).max() < 1e-6
        # instance 0 = the static-sphere problem
        static = BatchedMPC(desc, "f64").solve(b["x0"][:1], b["target"][:1], None if b["body_params"] is None else b["body_params"][:1])
        assert np.abs(out["X"][0][:, :27] - static["X"][0]).max() < 1e-8
    else:
        good = ok & (out["status"] == 0)
        assert good.sum() >= 4
        rx, ru = ranges(desc)
        ex = (np.abs(out["X"][good][:, :, :27] - ref["X"][good][:, :, :27]) / rx).reshape(good.sum(), -1).max(axis=1)
        assert np.median(ex) <= 2e-3 and ex.max() <= 2e-2
    # obstacle states of the solution = constant-acceleration prediction (full step) or the held guess (no step)
    t = dyn.dt * np.arange(dyn.N + 1)[None, :, None]
    pred = np.concatenate((xo[:, None, :3] + t * xo[:, None, 3:6] + 0.5 * t * t * xo[:, None, 6:],
                           xo[:, None, 3:6] + t * xo[:, None, 6:], np.tile(xo[:, None, 6:], (1, dyn.N + 1, 1))), axis=2)
    full = out["stats"][:, 3] == 1.0
    assert full.sum() >= 4 and np.abs(out["X"][full][:, :, 27:] - pred[full]).max() < (1e-9 if prec == "f64" else 1e-4)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_projectile_path_constraint_matches_oracle(prec):
    """ProjectilePathConstraint (projectile_path_constraint.h:46-156): probe values, warm- and cold-started solves
    with the flag s raised, and the flag lowered through ub_set_option, against the oracle (which carries the
    projectile as genuine states with the row's obstacle block - w s n' [I, t I, t^2/2 I])."""
    from _util import ballistic_prediction, projectile_problem, projectile_throws
    d, meta, tray = projectile_problem()
    d0, _, _ = projectile_problem(active=0.0)
    mpc = BatchedMPC(d, prec)
    assert mpc.nx == 36 and mpc.nx_robot == 27
    x0r = np.array(meta["x0"], dtype=float)
    throws = projectile_throws(oracle.fk(d, np.concatenate((x0r, np.zeros(9))))["spheres"][tray])
    Bn = len(throws)
    x0 = np.hstack((np.tile(x0r, (Bn, 1)), throws))
    target = np.tile(meta["r_ee0"] + np.array([0.1, 0.1, 0.05]), (Bn, d.N + 1, 1))
    h = mpc.eval("projectile_constraint", x0, np.zeros((Bn, mpc.nu)))
    href = np.array([oracle.projectile(d, x0[i])["h"] for i in range(Bn)])
    assert h.shape == (Bn, 2) and np.allclose(h, href, atol=1e-9)
    Xw = np.stack([np.hstack((np.tile(x0r, (d.N + 1, 1)), ballistic_prediction(xo, d.N, d.dt))) for xo in throws])
    Uw = np.zeros((Bn, d.N, mpc.nu))
    rx, ru = ranges(d)

    def check(out, ref):
        assert (ref["status"] == 0).all()
        if prec == "f64":
            assert (out["status"] == ref["status"]).all()
            assert np.abs(out["X"] - ref["X"]).max() < 1e-7 and np.abs(out["U"] - ref["U"]).max() < 1e-6
        else:
            assert (out["status"] == 0).sum() >= Bn - 1
            good = out["status"] == 0
            ex = (np.abs(out["X"][good][:, :, :27] - ref["X"][good][:, :, :27]) / rx).reshape(good.sum(), -1).max(axis=1)
            eu = (np.abs(out["U"][good] - ref["U"][good]) / ru).reshape(good.sum(), -1).max(axis=1)
            assert np.median(ex) <= 2e-3 and ex.max() <= 2e-2 and np.median(eu) <= 2e-3 and eu.max() <= 2e-2

    ref = oracle.solve_batch(d, x0, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    out = mpc.solve(x0, target, None, X=Xw.copy(), U=Uw.copy(), warm=True)
    check(out, ref)
    # the rows did something: the tray gives way where the flag-down plan does not
    ref0 = oracle.solve_batch(d0, x0, target, X=Xw.copy(), U=Uw.copy(), warm=True)
    assert np.abs(ref["X"][:, :, :27] - ref0["X"][:, :, :27]).max() > 1e-2
    if prec == "f64":   # cold start: obstacle states held over the horizon, rows linearised far from the prediction
        check(mpc.solve(x0, target), oracle.solve_batch(d, x0, target))
    mpc.set_option("projectile_active", 0)
    try:
        assert np.abs(mpc.eval("projectile_constraint", x0, np.zeros((Bn, mpc.nu)))).max() == 0
        check(mpc.solve(x0, target, None, X=Xw.copy(), U=Uw.copy(), warm=True), ref0)
    finally:
        mpc.set_option("projectile_active", 1)
    check(mpc.solve(x0, target, None, X=Xw.copy(), U=Uw.copy(), warm=True), ref)


def test_host_rollout_with_projectile_matches_oracle_engine():
    """The closed loop with the simulated projectile and the in-flight gate (rollout_host, mpc_sim.py:118-160 with
    mrt_node.cpp:241-263) on the fp64 kernels against the same loop on the CPU oracle: 14 warm-started replans
    with the flag going up and down reproduce the oracle's closed loop."""
    from _util import OracleEngine, projectile_rollout
    ref, desc, _ = projectile_rollout(OracleEngine, True)
    out, _, info = projectile_rollout(lambda d: BatchedMPC(d, "f64"), True)
    assert out["n_replans"] == ref["n_replans"] == 14 and (out["flags"] == ref["flags"]).all()
    assert np.abs(out["xs"] - ref["xs"]).max() < 1e-6 and np.abs(out["us"] - ref["us"]).max() < 1e-4
    assert np.abs(out["x_final"] - ref["x_final"]).max() < 1e-6


def test_operating_point_initializer_matches_oracle():
    """use_operating_points (controller_interface.cpp:380-387): the first solve starts from the operating trajectory
    instead of the held state — same result as the oracle started from that iterate, and not the cold-start one."""
    from upright_b200 import settings
    from upright_b200.manager import _RecedingHorizon
    from upright_b200.settings import TargetTrajectories
    from upright_b200.trajectory import StateInputTrajectory
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    x0 = np.array(meta["x0"], dtype=float)
    # operating trajectory: the arm eases 0.2 rad on joints 3..8 over 2 s with consistent velocities
    ts = np.linspace(0.0, 2.0, 11)
    prof, dprof = 0.5 * (1 - np.cos(np.pi * ts / 2.0)), 0.25 * np.pi * np.sin(np.pi * ts / 2.0)
    xs = np.tile(x0, (len(ts), 1))
    xs[:, 3:9] += 0.2 * prof[:, None]
    xs[:, 12:18] = 0.2 * dprof[:, None]
    us = np.zeros((len(ts), desc.nu))
    us[:, 9:] = 0.827 * 9.81 / 4                       # the forces that carry the object at rest
    st = settings.ControllerSettings(meta["controller_config"], x0=x0, operating_trajectory=StateInputTrajectory(ts, xs, us))
    st.use_operating_points = True
    mpc = BatchedMPC(st.to_desc(), "f64")
    rh = _RecedingHorizon(mpc, st, 2)
    r0 = mpc.eval("end_effector_position", x0, np.zeros(desc.nu))[0]
    goal = TargetTrajectories([0.0], [np.r_[r0 + [0.2, 0.1, 0.1], 0, 0, 0, 1, 0]], [np.zeros(desc.nu)])
    rh.reset([goal])
    rh.observe(0.0, np.tile(x0, (2, 1)))
    rh.advance()
    Xg, Ug = rh.operating_guess(0.0)
    target = np.tile(r0 + [0.2, 0.1, 0.1], (2, desc.N + 1, 1))
    ref = oracle.solve_batch(st.to_desc(), np.tile(x0, (2, 1)), target, X=Xg.copy(), U=Ug.copy(), warm=True)
    assert (rh.status == ref["status"]).all()
    assert np.abs(rh.X - ref["X"]).max() < 1e-7 and np.abs(rh.U - ref["U"]).max() < 1e-6
    cold = oracle.solve_batch(st.to_desc(), np.tile(x0, (2, 1)), target)
    assert np.abs(cold["X"] - ref["X"]).max() > 1e-3


@pytest.mark.parametrize("name", ["cfg3_thing_box_arch", "cfg5_thing_robust8", "cfg2_thing_demo"])
def test_run_time_dimension_kernels_match_oracle(name, monkeypatch):
    """UB_FORCE_GENERIC: the run-time-dimension kernels (a team of four warps per instance above 64 stage
    variables, one warp below) solve the BASELINE configurations like the specialised ones."""
    mpc, desc, meta = engine(name, "f64")
    b = batch_for(name, 4, 31)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    monkeypatch.setenv("UB_FORCE_GENERIC", "1")
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    monkeypatch.delenv("UB_FORCE_GENERIC")
    assert (out["status"] == ref["status"]).all()
    assert np.abs(out["X"] - ref["X"]).max() < 1e-7 and np.abs(out["U"] - ref["U"]).max() < 1e-6


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_device_closed_loop_with_dynamic_obstacle_matches_host_rollout(prec):
    """SURVEY.md §8(f) rank 2 on the device: `ub_closed_loop` carries the obstacle columns through the warm-start
    shift and the rollout and integrates the simulated obstacle with its mode schedule (free flight, then the sudden
    jump of obstacles/sudden.yaml) — against `rollout_host` with `plant.BallisticObstacles` driving the same kernels
    one `step(t, x)` at a time."""
    import copy
    from upright_b200.manager import BatchedControllerManager
    from upright_b200.plant import BallisticObstacles
    from upright_b200.settings import ControllerSettings, TargetTrajectories
    d4, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    cfg = copy.deepcopy(meta["controller_config"])
    x0r = np.array(meta["x0"], dtype=float)
    r_ee = np.array(meta["r_ee0"], dtype=float)
    start, jump = r_ee + [1.6, 0.3, 0.0], r_ee + [0.7, 0.1, 0.0]
    cfg["obstacles"]["dynamic"] = [{"name": "chair1", "radius": 0.25,
                                    "modes": [{"time": 0, "position": list(start), "velocity": [0, 0, 0], "acceleration": [0, 0, 0]}]}]
    cfg["obstacles"]["collision_pairs"] = list(cfg["obstacles"]["collision_pairs"]) + [["balanced_object_collision_link_0", "chair1"]]
    sim = {"controlled": False, "radius": 0.25, "relative": False,
           "modes": [{"time": 0, "position": list(start), "velocity": [-0.4, 0.0, 0.0], "acceleration": [0, 0, 0]},
                     {"time": 0.12, "position": list(jump), "velocity": [0, 0, 0], "acceleration": [0, 0, 0]}]}
    B, sim_dt, duration = 3, 0.01, 0.3
    rng = np.random.default_rng(5)
    x0 = np.tile(x0r, (B, 1))
    x0[:, :3] += rng.uniform(-0.05, 0.05, (B, 3))
    st = ControllerSettings(cfg)
    st.sqp.use_feedback_policy = False
    targets = [TargetTrajectories([0.0], [np.r_[r_ee + [0.1, 0.05, 0.0], 0, 0, 0, 1, 0]], [np.zeros(13)]) for _ in range(B)]
    plant = BallisticObstacles([sim], B)
    host = BatchedControllerManager(st, targets, timestep=0.05, precision=prec)
    ref = host.rollout_host(x0, duration, sim_dt, obstacles=plant)
    dev = BatchedControllerManager(st, targets, timestep=0.05, precision=prec)
    x0_full = np.hstack((x0, BallisticObstacles([sim], B).state()))
    out = dev.rollout(x0_full, duration, sim_dt, obstacles=[sim["modes"]])
    assert out["n_replans"] == ref["n_replans"] == 6
    assert out["xs"].shape == ref["xs"].shape == (B, 30, 36)
    tol = 1e-9 if prec == "f64" else 2e-3
    # the simulated obstacle: free flight, jump at the first step whose start time has reached 0.12 s
    assert np.abs(out["xs"][:, :, 27:] - ref["xs"][:, :, 27:]).max() < (1e-12 if prec == "f64" else 1e-5)
    assert np.abs(out["xs"][:, 13, 27:30] - jump).max() < 1e-5 and np.abs(out["xs"][:, 11, 27] - (start[0] - 0.4 * 0.11)).max() < 1e-5
    assert np.abs(out["xs"][:, :, :27] - ref["xs"][:, :, :27]).max() < tol
    assert np.abs(out["us"] - ref["us"]).max() < (1e-7 if prec == "f64" else 2e-2 * np.abs(ref["us"]).max())
    assert np.abs(out["x_final"] - ref["x_final"]).max() < tol
    # the jump matters: the plan after it differs from the plan of an obstacle that keeps flying
    far = BatchedControllerManager(st, targets, timestep=0.05, precision=prec)
    keep = far.rollout(x0_full, duration, sim_dt, obstacles=[sim["modes"][:1]])
    assert np.abs(keep["x_final"][:, :27] - out["x_final"][:, :27]).max() > 1e-4


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_end_effector_orientation_cost_matches_oracle(prec):
    """EndEffectorCost with a non-zero orientation weight (end_effector_cost.h:48-84: e = [r - r_d; quaternion
    distance], Gauss-Newton Hessian): kernels against the oracle on 7-column targets (position + desired quaternion),
    balancing on (soft rows), desired orientations tilted and turned away from the start."""
    import copy
    from upright_b200 import geometry as geo
    d0, meta = problem_io.load_fixture("cfg2_thing_demo")
    d = copy.deepcopy(d0)
    d.ee_weight[3], d.ee_weight[4], d.ee_weight[5] = 1.0, 2.0, 0.5
    b = batch_for("cfg2_thing_demo", 8, 23)
    mpc = BatchedMPC(d, prec)
    assert mpc.target_stride == 7
    rng = np.random.default_rng(9)
    tg = np.empty((8, d.N + 1, 7))
    for i in range(8):
        k = oracle.fk(d, b["x0"][i])
        q0 = geo.rot_to_quat(np.array(k["C"]).reshape(3, 3))
        ax = rng.standard_normal(3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(0.1, 0.5)
        qd = geo.quat_multiply(np.r_[np.sin(ang / 2) * ax, np.cos(ang / 2)], q0)
        tg[i, :, :3] = b["target"][i]
        tg[i, :, 3:] = qd
    ref = oracle.solve_batch(d, b["x0"], tg, b["body_params"])
    out = mpc.solve(b["x0"], tg, b["body_params"])
    plain = oracle.solve_batch(d0, b["x0"], b["target"], b["body_params"])
    assert (ref["status"] == 0).all() and (out["status"] == ref["status"]).all()
    assert np.abs(ref["X"] - plain["X"]).max() > 1e-2                    # the orientation term acts
    rx, ru = ranges(d)
    ex = (np.abs(out["X"] - ref["X"]) / rx).max()
    eu = (np.abs(out["U"] - ref["U"]) / ru).max()
    print(f"orientation cost {prec}: scaled error X {ex:.2e} U {eu:.2e}")
    if prec == "f64":
        assert (out["stats"][:, 0] == ref["stats"][:, 0]).all() and np.abs(out["X"] - ref["X"]).max() < 1e-7
        assert np.abs(out["stats"][:, 1] - ref["stats"][:, 1]).max() < 1e-8   # cost incl. the orientation term
    else:
        assert ex < 1e-3 and eu < 1e-3


def _ground_desc():
    """cfg4 plus the reference's `ground` half-space paired with the wrist sphere (obstacles/dynamic.yaml:28-29),
    raised to z <= 0.41: the wrist sphere (z = 0.69 at the start, radius 0.15, minimum distance 0.1) then starts 3 cm
    inside the feasible side and the goals that ask the tray down by up to 8 cm make the row active."""
    import copy
    from upright_b200 import bindings as Bd
    d0, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    d = copy.deepcopy(d0)
    robot_slots = [i for i in range(d.n_spheres) if d.spheres[i].link >= 0]
    wrist = min(robot_slots, key=lambda i: abs(d.spheres[i].link - (d.nq - 1)))
    g = d.n_spheres
    d.spheres[g].link, d.spheres[g].shape, d.spheres[g].radius = -1, Bd.UB_SHAPE_HALFSPACE, 0.41
    d.spheres[g].offset[:] = [0.0, 0.0, 1.0]
    d.n_spheres += 1
    d.pairs[d.n_pairs].a, d.pairs[d.n_pairs].b = wrist, g
    d.n_pairs += 1
    return d0, d, meta


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_ground_half_space_pair_matches_oracle(prec):
    """Sphere / half-space rows (add_ground_plane, controller_interface.cpp:93-101,189): probe values and full solves
    (run-time-dimension kernel: 13 pair rows) against the oracle."""
    d0, d, meta = _ground_desc()
    b = batch_for("cfg4_thing_obstacles2", 8, 5)
    mpc = BatchedMPC(d, prec)
    h = mpc.eval("obstacle_avoidance", b["x0"], np.zeros((8, mpc.nu)))
    href = np.array([oracle.linearize(d, b["x0"][i], np.zeros(mpc.nu))["hobs"] for i in range(8)])
    assert h.shape == (8, 13) and np.allclose(h, href, atol=1e-10)
    ref = oracle.solve_batch(d, b["x0"], b["target"], b["body_params"])
    free = oracle.solve_batch(d0, b["x0"], b["target"], b["body_params"])
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    ok = ref["status"] == 0
    assert ok.sum() >= 5 and (out["status"][ok] == 0).all()
    rx, ru = ranges(d)
    ex = (np.abs(out["X"][ok] - ref["X"][ok]) / rx).max()
    eu = (np.abs(out["U"][ok] - ref["U"][ok]) / ru).max()
    print(f"ground pair {prec}: scaled error X {ex:.2e} U {eu:.2e}; moved by the row {np.abs(ref['X'] - free['X']).max():.2e}")
    assert ex < (1e-7 if prec == "f64" else 1e-3) and eu < (1e-6 if prec == "f64" else 1e-3)
    assert np.abs(ref["X"] - free["X"]).max() > 1e-3          # the row acts


def test_model_queries_of_the_python_interface():
    """flowMap, flowMapLinearApproximation, costQuadraticApproximation (pybindings.cpp:388-397) and
    BalancingConstraintWrapper.getLinearApproximation (balancing_constraint_wrapper.h:45-60) on the shim, against
    the oracle's linearisation and finite differences."""
    from upright_b200.manager import BalancingConstraintWrapper, ControllerInterface
    from upright_b200.settings import ControllerSettings, TargetTrajectories
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    st = ControllerSettings(config=meta["controller_config"], x0=np.array(meta["x0"]))
    ci = ControllerInterface(st, precision="f64")
    rng = np.random.default_rng(3)
    x = np.array(meta["x0"], dtype=float) + 0.1 * rng.standard_normal(27)
    u = rng.standard_normal(13)
    r0 = ci._engine.eval("end_effector_position", x, u)[0]
    ci.reset(TargetTrajectories([0.0], [np.r_[r0 + [0.1, -0.2, 0.05], 0, 0, 0, 1, 0]], [np.zeros(13)]))
    f = ci.flowMap(0.0, x, u)
    assert np.allclose(f, np.r_[x[9:27], u[:9]])
    lin = ci.flowMapLinearApproximation(0.0, x, u)
    assert np.allclose(lin.dfdx @ x + lin.dfdu @ u, f) and lin.dfdx.shape == (27, 27) and lin.dfdu.shape == (27, 13)
    q = ci.costQuadraticApproximation(0.0, x, u)
    assert q.f == pytest.approx(ci.cost(0.0, x, u), rel=1e-12)
    gx = np.array([(ci.cost(0.0, x + e, u) - ci.cost(0.0, x - e, u)) / 2e-6 for e in 1e-6 * np.eye(27)])
    gu = np.array([(ci.cost(0.0, x, u + e) - ci.cost(0.0, x, u - e)) / 2e-6 for e in 1e-6 * np.eye(13)])
    assert np.allclose(q.dfdx, gx, atol=1e-6) and np.allclose(q.dfdu, gu, atol=1e-6)
    assert np.allclose(q.dfdxx, q.dfdxx.T) and np.linalg.eigvalsh(q.dfdxx).min() > -1e-12 and q.dfdux.shape == (13, 27)
    assert ci.getCostValue("state_input_cost", 0.0, x, u) + ci.getCostValue("end_effector_cost", 0.0, x, u) == pytest.approx(q.f)
    w = BalancingConstraintWrapper(st)
    a = w.getLinearApproximation(0.0, x, u)
    ol = oracle.linearize(desc, x, u)
    assert a.f.shape == (6,) and np.allclose(a.f, ol["g"], atol=1e-11)
    assert a.dfdx.shape == (6, 27) and np.allclose(a.dfdx, ol["C"], atol=1e-10) and np.allclose(a.dfdu_dynamics[:, 9:], ol["Df"], atol=1e-12)
    assert np.all(a.dfdu_dynamics[:, :9] == 0)
    # friction-cone configuration: contact rows first, dynamics rows behind (approx.f << a.f, b.f)
    d3, m3 = problem_io.load_fixture("cfg3_thing_box_arch")
    w3 = BalancingConstraintWrapper(ControllerSettings(config=m3["controller_config"], x0=np.array(m3["x0"])))
    u3 = rng.standard_normal(57)
    a3 = w3.getLinearApproximation(0.0, x, u3)
    o3 = oracle.linearize(d3, x, u3)
    assert a3.f.shape == (80 + 18,) and np.allclose(a3.f[:80], o3["hfric"], atol=1e-12) and np.allclose(a3.f[80:], o3["g"], atol=1e-10)
    assert np.all(a3.dfdx[:80] == 0) and np.allclose(a3.dfdx[80:], o3["C"], atol=1e-10)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE batch sizes (cfg1 / cfg3: 4096 on one GPU; cfg4: 2048 = one GPU's share of 16384; cfg5: 1024 of 8192)


@pytest.mark.parametrize("name,B", [("cfg1_ur10_demo", 4096), ("cfg3_thing_box_arch", 4096), ("cfg4_thing_obstacles2", 2048),
                                    ("cfg5_thing_robust8", 1024)])
def test_full_baseline_batches(name, B):
    """No instance is left without a finite answer (status NAN count 0, no fp64 rescue); accepted full steps are
    dynamically consistent; sampled instances agree with the oracle within the stated tolerance."""
    mpc, desc, meta = engine(name, "f32")
    b = batch_for(name, B, 4242)
    out = mpc.solve(b["x0"], b["target"], b["body_params"])
    assert (out["status"] != 3).all() and np.isfinite(out["X"]).all() and np.isfinite(out["U"]).all()
    assert (out["status"] == 0).mean() > 0.9
    X, U, nq, dt = out["X"], out["U"], desc.nq, desc.dt
    full = out["stats"][:, 3] == 1.0
    q, v, a, j = X[:, :-1, :nq], X[:, :-1, nq:2 * nq], X[:, :-1, 2 * nq:], U[:, :, :nq]
    gap = np.abs(q + dt * v + 0.5 * dt * dt * a + dt**3 / 6 * j - X[:, 1:, :nq]).max((1, 2))
    assert full.mean() > 0.8 and gap[full].max() < 5e-6       # float storage of the iterate
    sub = np.random.default_rng(1).choice(B, 12, replace=False)
    bp = None if b["body_params"] is None else b["body_params"][sub]
    ref = oracle.solve_batch(desc, b["x0"][sub], b["target"][sub], bp)
    assert (ref["status"] == out["status"][sub]).all()
    rx, ru = ranges(desc)
    ok = ref["status"] == 0
    assert ok.sum() >= 8
    assert (np.abs(X[sub][ok] - ref["X"][ok]) / rx).max() < 1e-3 and (np.abs(U[sub][ok] - ref["U"][ok]) / ru).max() < 1e-3


def test_device_closed_loop_keeps_the_object_balanced():
    """The physical check of tests/test_host_api.py on the product path: 3 s of `ub_closed_loop` for 32 robots with
    the fp32 kernels — non-negative normal forces that carry the object exist at every logged state."""
    from scipy.optimize import nnls
    from upright_b200.manager import BatchedControllerManager
    from upright_b200.settings import ControllerSettings, TargetTrajectories
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    st = ControllerSettings(config=meta["controller_config"], x0=np.array(meta["x0"]))
    B = 32
    rng = np.random.default_rng(11)
    x0 = np.tile(np.array(meta["x0"], dtype=float), (B, 1))
    x0[:, :4] += rng.uniform(-0.1, 0.1, (B, 4))       # base x, y, yaw and shoulder pan: the tray starts level
    probe = BatchedControllerManager(st, [None] * B)
    r0 = probe.engine.eval("end_effector_position", x0, np.zeros((B, 13)))
    targets = [TargetTrajectories([0.0], [np.r_[r0[b] + [-0.25, 0.5, 0.25], 0, 0, 0, 1, 0]], [np.zeros(13)]) for b in range(B)]
    mgr = BatchedControllerManager(st, targets, timestep=0.05)
    out = mgr.rollout(x0, 3.0, 0.01, log_stride=10)
    assert (out["status_counts"][:, 3] == 0).all()
    at_rest = 9.81 / np.sqrt(6.0)
    worst = 0.0
    for b in range(0, B, 4):
        for x in out["xs"][b]:
            lin = oracle.linearize(desc, x, np.zeros(13))
            worst = max(worst, nnls(lin["Df"], -lin["g"])[1])
    assert worst < 0.05 * at_rest


def test_robust_planner_shape_n100():
    """The planner shape of the robust demos (upright_robust/config/demos/_base.yaml:62-66: time_horizon 10 s at
    dt 0.1 -> N = 100 knots, init_sqp_iteration 3) on the robust 8-vertex constraint set (cfg5): fp64 kernels against
    the oracle at the fp64 tolerance, the product kernels at the stated fp32 tolerance.  Runs the run-time-dimension
    kernel (the specialised ones are built for N = 20)."""
    import copy
    base, meta = problem_io.load_fixture("cfg5_thing_robust8")
    desc = copy.deepcopy(base)
    desc.N, desc.sqp_iteration = 100, 3
    B = 8
    ee = lambda x: np.stack([oracle.fk(desc, xi)["r"] for xi in x])  # noqa: E731
    b = workload.sample_batch("cfg5_thing_robust8", desc, meta, B, 11, ee)
    assert b["target"].shape == (B, 101, 3)
    ref = oracle.solve_batch(desc, b["x0"], b["target"], b["body_params"])
    assert (ref["status"] == 0).all() and (ref["stats"][:, 7] == 3).all()
    rx, ru = ranges(desc)
    for prec, tol in (("f64", 1e-7), ("f32", 1e-3)):
        mpc = BatchedMPC(desc, prec)
        assert mpc.N == 100
        out = mpc.solve(b["x0"], b["target"], b["body_params"])
        assert (out["status"] == 0).all(), out["status"]
        assert np.array_equal(out["stats"][:, 7], ref["stats"][:, 7])           # three SQP iterations each
        ex = (np.abs(out["X"] - ref["X"]) / rx).max()
        eu = (np.abs(out["U"] - ref["U"]) / ru).max()
        print(f"N=100 robust planner {prec}: scaled error X {ex:.2e} U {eu:.2e}; interior-point iterations "
              f"{out['stats'][:, 0].tolist()} (oracle {ref['stats'][:, 0].tolist()})")
        if prec == "f64":
            assert np.array_equal(out["stats"][:, 0], ref["stats"][:, 0])
            assert np.abs(out["X"] - ref["X"]).max() <= tol and np.abs(out["U"] - ref["U"]).max() <= tol
        else:
            assert max(ex, eu) <= tol


def test_phase_alignment_does_not_change_results():
    """The meetings that keep the warps of a CTA in step (ub::CtaAlign; on by default from two waves of the persistent
    grid) only order the work in time: forced on and forced off give bitwise identical trajectories, on a batch of
    several waves with mixed iteration counts and with the batch tail (warps out of work that keep attending)."""
    import os
    name = "cfg2_thing_demo"
    mpc, desc, meta = engine(name, "f32")
    B = 6100                                   # 2.6 waves of the 2368 resident warps of a B200, not a multiple of anything
    b = batch_for(name, B, 31)
    dev = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")  # noqa: E731
    x0, tg, bp = dev(b["x0"]), dev(b["target"]), dev(b["body_params"])
    outs = {}
    old = os.environ.get("UB_ALIGN_GROUP")
    try:
        for mode in ("0", "16", "4"):
            os.environ["UB_ALIGN_GROUP"] = mode
            o = mpc.solve_device(x0, tg, bp)
            torch.cuda.synchronize()
            outs[mode] = {k: o[k].clone() for k in ("X", "U", "status")}
    finally:
        if old is None:
            os.environ.pop("UB_ALIGN_GROUP", None)
        else:
            os.environ["UB_ALIGN_GROUP"] = old
    o = mpc.solve_device(x0, tg, bp)           # default: on for this batch size
    torch.cuda.synchronize()
    assert (outs["0"]["status"] == 0).float().mean() > 0.98
    for mode in ("16", "4"):
        for k in ("X", "U", "status"):
            assert torch.equal(outs["0"][k], outs[mode][k]), (mode, k)
    assert torch.equal(outs["0"]["X"], o["X"]) and torch.equal(outs["0"]["U"], o["U"])


def test_wrench_cone_verification_of_planned_trajectories():
    """SURVEY.md §8f rank 4 (process_sim_runs.py:208-246, exact parameters): the contact-wrench-cone face form of the
    arrangement checked at every knot of a batch of PLANNED trajectories — wrenches from the probe kernel, the check one
    GEMM on the GPU — against the same quantity computed knot by knot with the oracle; planned motions stay inside the
    cone up to the slack of the soft rows, an unplanned violent motion does not."""
    from upright_b200.robust import WrenchConeVerifier
    name = "cfg2_thing_demo"
    mpc, desc, meta = engine(name, "f32")
    B = 64
    b = workload.sample_batch(name, desc, meta, B, 5, lambda x: mpc.eval("end_effector_position", x, np.zeros((x.shape[0], mpc.nu))),
                              vary_bodies=False)
    out = mpc.solve(b["x0"], b["target"], None)
    ver = WrenchConeVerifier(mpc, desc)
    V = ver.violation(out["X"][:, :-1])                     # the rows exist at the intermediate knots
    assert V.shape == (B, desc.N)
    # parity with the CPU evaluation of the same face form
    for i in (0, 17, 63):
        for k in (0, 7, 19):
            w = -oracle.linearize(desc, out["X"][i, k], np.zeros(desc.nu))["g"]
            assert abs(float((ver.A @ w).max()) - V[i, k]) < 1e-9
    # Planned knots (k >= 1; knot 0 is the observation: the seeded start states tilt the tray, which frictionless
    # contacts cannot carry) approach the cone as the SQP iterations converge (gravity row of the wrench: 4.0)
    med = [np.median(V[:, 1:].max(1))]
    for _ in range(3):
        out = mpc.solve(b["x0"], b["target"], None, X=out["X"], U=out["U"], warm=True)
        med.append(np.median(ver.violation(out["X"][:, :-1])[:, 1:].max(1)))
    print("wrench cone: median over instances of the largest planned violation, per SQP iteration:", ["%.3e" % m for m in med])
    assert med[1] < 0.5 * med[0] and med[2] < 0.5 * med[1] and med[3] < 0.05
    # an unplanned motion: 1 g of base acceleration at every knot
    Xbad = out["X"][:4, :-1].copy()
    Xbad[:, :, 2 * desc.nq:] = 0.0
    Xbad[:, :, 2 * desc.nq] = 9.81
    Vbad = ver.violation(Xbad)
    print(f"wrench cone: violent motion median {np.median(Vbad):.3f} min {Vbad.min():.3f}")
    assert np.median(Vbad) > 1.0 and (Vbad > 0.2).mean() > 0.9
