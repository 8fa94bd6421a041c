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
