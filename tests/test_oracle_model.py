"""Oracle model functions: analytic sanity cases and finite-difference checks
of every Jacobian the QP uses (constraint values, FK and
Jacobians are pinned by no reference test, so the oracle is cross-checked
against first principles)."""
import numpy as np
import pytest

from upright_b200 import problem_io

CFGS = list(problem_io.FIXTURES)


def _state(desc, meta, rng, scale=0.3):
    nq = desc.nq
    x = np.array(meta["x0"], dtype=float)
    x[:nq] += scale * rng.standard_normal(nq)
    x[nq:] = scale * rng.standard_normal(2 * nq)
    return x


@pytest.mark.parametrize("name", ["cfg1_ur10_demo", "cfg2_thing_demo"])
def test_home_pose_is_level(oracle_lib, name):
    desc, meta = problem_io.load_fixture(name)
    k = oracle_lib.fk(desc, meta["x0"])
    assert np.allclose(k["C"], np.eye(3), atol=1e-7)          # tray level, axes world-aligned
    assert np.allclose(k["r"], meta["r_ee0"], atol=1e-12)     # numpy host FK == oracle FK
    assert np.allclose(k["v"], 0) and np.allclose(k["w"], 0) and np.allclose(k["a"], 0)


def test_fixed_base_equals_locked_thing(oracle_lib):
    d6, m6 = problem_io.load_fixture("cfg1_ur10_demo")
    d9, m9 = problem_io.load_fixture("cfg2_thing_demo")
    rng = np.random.default_rng(3)
    x6 = _state(d6, m6, rng)
    x9 = np.zeros(27)
    x9[:3] = [-1.0, 1.0, 0.0]  # ur10.yaml:54 base_pose
    for blk in range(3):
        x9[9 * blk + 3: 9 * blk + 9] = x6[6 * blk: 6 * blk + 6]
    k6, k9 = oracle_lib.fk(d6, x6), oracle_lib.fk(d9, x9)
    for key in ("r", "C", "v", "w", "a", "alpha"):
        assert np.allclose(k6[key], k9[key], atol=1e-12), key


def test_static_balance_is_feasible(oracle_lib):
    """Level tray at rest with f_i = m g / 4 through symmetric contacts: all six
    object-dynamics rows vanish."""
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    u = np.zeros(desc.nu)
    u[desc.nq:] = 0.827 * 9.81 / 4
    lin = oracle_lib.linearize(desc, meta["x0"], u)
    assert np.abs(lin["g"]).max() < 1e-7  # calibration rpy is rounded to 8 digits in the YAML


@pytest.mark.parametrize("name", CFGS)
def test_velocity_and_acceleration_consistency(oracle_lib, name):
    desc, meta = problem_io.load_fixture(name)
    rng = np.random.default_rng(1)
    nq = desc.nq
    x = _state(desc, meta, rng)
    k = oracle_lib.fk(desc, x)
    lin = oracle_lib.linearize(desc, x, np.zeros(desc.nu))
    assert np.allclose(k["v"], lin["Jp"] @ x[nq:2 * nq], atol=1e-12)

    def along(t):
        return np.concatenate([x[:nq] + x[nq:2 * nq] * t + 0.5 * x[2 * nq:] * t * t, x[nq:2 * nq] + x[2 * nq:] * t, x[2 * nq:]])
    h = 1e-5
    kp, km = oracle_lib.fk(desc, along(h)), oracle_lib.fk(desc, along(-h))
    assert np.allclose((kp["v"] - km["v"]) / (2 * h), k["a"], atol=1e-7)       # classical acceleration
    assert np.allclose((kp["w"] - km["w"]) / (2 * h), k["alpha"], atol=1e-7)
    W = ((kp["C"] - km["C"]) / (2 * h)) @ k["C"].T                              # C' = skew(w) C
    assert np.allclose([W[2, 1], W[0, 2], W[1, 0]], k["w"], atol=1e-7)


@pytest.mark.parametrize("name", CFGS)
def test_jacobians_finite_difference(oracle_lib, name):
    desc, meta = problem_io.load_fixture(name)
    rng = np.random.default_rng(2)
    nq, nx, nu = desc.nq, desc.nx, desc.nu
    x = _state(desc, meta, rng)
    u = rng.standard_normal(nu)
    lin = oracle_lib.linearize(desc, x, u)
    eps = 1e-6

    def fd(fun, n, idx):
        J = np.zeros((n, len(idx)))
        for c, j in enumerate(idx):
            xp, xm = x.copy(), x.copy()
            xp[j] += eps
            xm[j] -= eps
            J[:, c] = (fun(xp) - fun(xm)) / (2 * eps)
        return J
    assert np.allclose(fd(lambda z: oracle_lib.fk(desc, z)["r"], 3, range(nq)), lin["Jp"], atol=1e-8)
    if desc.n_eq:
        Cfd = fd(lambda z: oracle_lib.linearize(desc, z, u)["g"], desc.n_eq, range(nx))
        assert np.allclose(Cfd, lin["C"], atol=2e-7)
        Dfd = np.zeros_like(lin["Df"])
        for j in range(nu - nq):
            up, um = u.copy(), u.copy()
            up[nq + j] += 1.0
            um[nq + j] -= 1.0
            Dfd[:, j] = (oracle_lib.linearize(desc, x, up)["g"] - oracle_lib.linearize(desc, x, um)["g"]) / 2
        assert np.allclose(Dfd, lin["Df"], atol=1e-10)
    if desc.n_obs:
        Jfd = fd(lambda z: oracle_lib.linearize(desc, z, u)["hobs"], desc.n_obs, range(nq))
        assert np.allclose(Jfd, lin["Jobs"], atol=1e-7)


def test_friction_rows_match_formula(oracle_lib):
    """compute_contact_force_constraints_linearized (contact_constraints.h:49-77)."""
    desc, meta = problem_io.load_fixture("cfg3_thing_box_arch")
    rng = np.random.default_rng(5)
    u = rng.standard_normal(desc.nu)
    lin = oracle_lib.linearize(desc, meta["x0"], u)
    f = u[desc.nq:].reshape(-1, 3)
    for i in range(desc.nc):
        c = desc.contacts[i]
        n, S = np.array(list(c.normal)), np.array(list(c.span)).reshape(2, 3)
        fn, ft = n @ f[i], S @ f[i]
        exp = [fn, c.mu * fn - ft[0] - ft[1], c.mu * fn - ft[0] + ft[1], c.mu * fn + ft[0] - ft[1], c.mu * fn + ft[0] + ft[1]]
        assert np.allclose(lin["hfric"][5 * i: 5 * i + 5], exp, atol=1e-12)
    assert np.allclose(lin["Ffric"] @ u[desc.nq:], lin["hfric"], atol=1e-12)


def test_object_dynamics_against_direct_numpy(oracle_lib):
    """Independent numpy evaluation of contact_constraints.h:79-157 for one body."""
    from upright_b200.geometry import skew3
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    rng = np.random.default_rng(9)
    x = _state(desc, meta, rng)
    u = rng.standard_normal(desc.nu)
    k = oracle_lib.fk(desc, x)
    p = np.array([desc.body_params[0][j] for j in range(10)])
    m, com = p[0], p[1:4] / p[0]
    I = np.array([[p[4], p[5], p[6]], [p[5], p[7], p[8]], [p[6], p[8], p[9]]])
    C, w, al, a = k["C"], k["w"], k["alpha"], k["a"]
    ddC = (skew3(al) + skew3(w) @ skew3(w)) @ C
    gi = m * C.T @ (a + ddC @ com - np.array(list(desc.gravity)))
    we, ale = C.T @ w, C.T @ al
    tau = np.cross(we, I @ we) + I @ ale
    F, Tq = np.zeros(3), np.zeros(3)
    for i in range(desc.nc):
        c = desc.contacts[i]
        f = u[desc.nq + i] * np.array(list(c.normal))
        F -= f
        Tq += np.cross(np.array(list(c.r_co_o2)) - com, -f)
    exp = np.concatenate([(gi - F) / m, (tau - Tq) / m]) / np.sqrt(6.0)
    assert np.allclose(oracle_lib.linearize(desc, x, u)["g"], exp, atol=1e-12)


def test_ground_half_space_row_in_the_oracle(oracle_lib):
    import copy
    from upright_b200 import bindings as B
    oracle = oracle_lib
    """Sphere against the half-space z <= 0: value c_z - r - minimum_distance, Jacobian = z-row of the sphere's
    position Jacobian (central differences), and the row keeps the wrist above the floor in a solve."""
    d0, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    d = copy.deepcopy(d0)
    robot_slots = [i for i in range(d.n_spheres) if d.spheres[i].link >= 0]
    g = d.n_spheres
    d.spheres[g].link, d.spheres[g].shape, d.spheres[g].radius = -1, B.UB_SHAPE_HALFSPACE, 0.0
    d.spheres[g].offset[:] = [0.0, 0.0, 1.0]
    d.n_spheres += 1
    a = robot_slots[0]
    d.pairs[d.n_pairs].a, d.pairs[d.n_pairs].b = a, g
    d.n_pairs += 1
    x = np.array(meta["x0"], dtype=float)
    x[:9] += 0.1 * np.random.default_rng(0).standard_normal(9)
    nu = oracle.dims(d)["nu"]
    lin = oracle.linearize(d, x, np.zeros(nu))
    c = oracle.fk(d, x)["spheres"][a]
    assert lin["hobs"][-1] == pytest.approx(c[2] - d.spheres[a].radius - d.minimum_distance, abs=1e-12)
    J = np.zeros(9)
    for j in range(9):
        e = np.zeros(27)
        e[j] = 1e-6
        J[j] = (oracle.linearize(d, x + e, np.zeros(nu))["hobs"][-1] - oracle.linearize(d, x - e, np.zeros(nu))["hobs"][-1]) / 2e-6
    assert np.allclose(lin["Jobs"][-1], J, atol=1e-7)
    assert np.allclose(lin["hobs"][:-1], oracle.linearize(d0, x, np.zeros(nu))["hobs"], atol=1e-14)   # other rows untouched


def test_end_effector_orientation_error(oracle_lib):
    """ocs2 quaternionDistance of the measured against the desired quaternion (end_effector_cost.h:61-67 [EXT]):
    zero at the target, sin(angle / 2) * axis for a desired orientation rotated by (axis, angle) in the world frame,
    Jacobian = central differences; and the cost acts in a solve (7-column targets)."""
    import copy
    from upright_b200 import geometry as geo
    oracle = oracle_lib
    d0, meta = problem_io.load_fixture("cfg2_thing_demo")
    d = copy.deepcopy(d0)
    d.ee_weight[3] = d.ee_weight[4] = d.ee_weight[5] = 2.0
    assert oracle.target_stride(d0) == 3 and oracle.target_stride(d) == 7
    rng = np.random.default_rng(0)
    x = np.array(meta["x0"], dtype=float)
    x[:9] += 0.3 * rng.standard_normal(9)
    k = oracle.fk(d, x)
    q = geo.rot_to_quat(np.array(k["C"]).reshape(3, 3))
    assert np.abs(oracle.orientation_error(d, x, q)[0]).max() < 1e-15
    ang, ax = 0.2, np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    qref = geo.quat_multiply(np.r_[np.sin(ang / 2) * ax, np.cos(ang / 2)], q)
    e, J = oracle.orientation_error(d, x, qref)
    assert np.allclose(e, np.sin(ang / 2) * ax, atol=1e-12)
    Jn = np.zeros((3, 9))
    for j in range(9):
        dx = np.zeros(27)
        dx[j] = 1e-6
        Jn[:, j] = (oracle.orientation_error(d, x + dx, qref)[0] - oracle.orientation_error(d, x - dx, qref)[0]) / 2e-6
    assert np.allclose(J, Jn, atol=1e-8)
    # solves: the weighted problem turns the tray towards the desired orientation, the unweighted one does not care
    # (balancing off: with it the tilt is dictated by the object)
    for dd in (d, d0):
        dd.balancing_enabled = 0
    N = d.N
    x0 = np.array(meta["x0"], dtype=float)
    k0 = oracle.fk(d, x0)
    q0 = geo.rot_to_quat(np.array(k0["C"]).reshape(3, 3))
    qd = geo.quat_multiply(np.r_[0, 0, np.sin(0.15), np.cos(0.15)], q0)          # 0.3 rad about the world z axis
    tg7 = np.tile(np.r_[k0["r"], qd], (1, N + 1, 1))
    out = oracle.solve_batch(d, x0[None], tg7)
    ref = oracle.solve_batch(d0, x0[None], tg7[:, :, :3])
    assert out["status"][0] == 0 and ref["status"][0] == 0
    e_w = np.linalg.norm(oracle.orientation_error(d, out["X"][0, -1], qd)[0])
    e_0 = np.linalg.norm(oracle.orientation_error(d, ref["X"][0, -1], qd)[0])
    assert e_0 == pytest.approx(np.sin(0.15), abs=1e-6) and e_w < 0.5 * e_0
