"""Host-side logic that needs no GPU: settings parsing from the reference's own
YAML (when present) against the committed fixtures, target trajectories,
workload generator, receding-horizon bookkeeping with a stub engine."""
import copy
from pathlib import Path

import numpy as np
import pytest

from upright_b200 import bindings as B
from upright_b200 import config, problem_io, settings, workload
from upright_b200.settings import TargetTrajectories

REF = Path("/root/reference")


def test_fixture_dimensions_match_survey_table():
    expect = {  # SURVEY.md §8 dimension table: nq, nx, nb, nc, nf, nu, eq, N
        "cfg1_ur10_demo": (6, 18, 1, 4, 1, 10, 6, 20),
        "cfg2_thing_demo": (9, 27, 1, 4, 1, 13, 6, 20),
        "cfg3_thing_box_arch": (9, 27, 3, 16, 3, 57, 18, 20),
        "cfg4_thing_obstacles2": (9, 27, 1, 4, 1, 13, 6, 20),
        "cfg5_thing_robust8": (9, 27, 8, 32, 1, 41, 48, 20),
    }
    for name, e in expect.items():
        d, _ = problem_io.load_fixture(name)
        assert (d.nq, d.nx, d.nb, d.nc, d.nf, d.nu, d.n_eq, d.N) == e, name
    d3, _ = problem_io.load_fixture("cfg3_thing_box_arch")
    assert d3.n_fric == 80
    d4, _ = problem_io.load_fixture("cfg4_thing_obstacles2")
    assert d4.n_obs == 12 and d4.slacks.enabled == 0
    d5, _ = problem_io.load_fixture("cfg5_thing_robust8")
    assert d5.slacks.enabled == 1 and d5.slacks.input_box == 0 and d5.force_weight == 0.0


@pytest.mark.skipif(not REF.exists(), reason="reference tree only in the build container")
@pytest.mark.parametrize("name", list(problem_io.FIXTURES))
def test_fixtures_regenerate_from_reference_yaml(name):
    rel = problem_io.FIXTURES[name]
    pkg, path = rel.split("/", 1)
    cfg = config.load_config(config.resolve_package_path({"package": pkg, "path": path}))
    s = settings.ControllerSettings(cfg["controller"])
    fresh = problem_io.desc_to_dict(s.to_desc())
    stored = problem_io.desc_to_dict(problem_io.load_fixture(name)[0])

    def close(a, b):
        if isinstance(a, dict):
            return all(close(a[k], b[k]) for k in a)
        if isinstance(a, list):
            return all(close(x, y) for x, y in zip(a, b))
        return np.isclose(a, b, rtol=1e-12, atol=1e-12)
    assert close(fresh, stored)


@pytest.mark.skipif(not REF.exists(), reason="reference tree only in the build container")
def test_settings_surface_thing_demo():
    cfg = config.load_config(REF / "upright_cmd/config/demos/thing_demo.yaml")["controller"]
    s = settings.ControllerSettings(cfg)
    assert s.mpc.time_horizon == 2.0 and s.sqp.dt == 0.1 and s.sqp.sqp_iteration == 1
    assert s.sqp.hpipm.iter_max == 30 and s.sqp.hpipm.slacks.enabled is True
    assert s.sqp.hpipm.slacks.upper_L2_penalty == 100
    assert (s.dims.robot.q, s.dims.robot.x, s.dims.c, s.dims.nf) == (9, 27, 4, 1)
    assert s.dims.x() == 27 and s.dims.u() == 13 and s.dims.f() == 4
    assert np.allclose(np.diag(s.state_weight), 0.01 * np.r_[np.zeros(9), 10 * np.ones(9), np.ones(9)])
    assert np.allclose(np.diag(s.input_weight), 0.001)
    assert s.balancing_settings.force_weight == 0.001
    assert np.allclose(s.gravity, [0, 0, -9.81])
    assert np.allclose(s.initial_state[:9], [-1, 1, 0, 0.5 * np.pi, -0.25 * np.pi, 0.5 * np.pi, -0.25 * np.pi, 0.5 * np.pi, 0.417 * np.pi])
    assert s.tracking.min_policy_update_time == 0.01


def test_orientation_weight_reaches_the_description():
    """A non-zero orientation weight (end_effector_cost.h:61-81; zero in every shipped configuration) is carried into
    the problem description, and targets then hold the desired quaternion behind the position."""
    if REF.exists():
        cfg = config.load_config(REF / "upright_cmd/config/demos/thing_demo.yaml")["controller"]
        cfg["weights"]["end_effector"]["diag"] = [1, 1, 1, 0.5, 0.25, 0]
        d = settings.ControllerSettings(cfg).to_desc()
        assert list(d.ee_weight) == [1, 1, 1, 0.5, 0.25, 0]


def test_pose_targets_slerp_like_eigen():
    """interpolate_end_effector_pose (reference_trajectory.h:18-47): position linear, orientation
    q_lhs.slerp(1 - alpha, q_rhs); single waypoint constant; antipodal quaternions take the shortest arc."""
    from upright_b200 import geometry as geo
    from upright_b200.settings import quat_slerp
    q0 = np.array([0.0, 0.0, 0.0, 1.0])
    q1 = geo.rot_to_quat(geo.rotz(1.0))
    tt = TargetTrajectories([0.0, 2.0], [np.r_[0, 0, 0, q0, 0], np.r_[2, 0, 0, q1, 0]], [np.zeros(3)] * 2)
    p = tt.poses_at([-1.0, 0.0, 0.5, 2.0, 3.0])
    assert np.allclose(p[0], np.r_[0, 0, 0, q0]) and np.allclose(p[4], np.r_[2, 0, 0, q1])
    assert np.allclose(p[2][:3], [0.5, 0, 0]) and np.allclose(p[2][3:], geo.rot_to_quat(geo.rotz(0.25)), atol=1e-12)
    assert np.allclose(quat_slerp(q0, -q1, 0.5), -geo.rot_to_quat(geo.rotz(0.5)) * np.sign(1.0), atol=1e-12) or \
        np.allclose(quat_slerp(q0, -q1, 0.5), geo.rot_to_quat(geo.rotz(0.5)), atol=1e-12)
    assert np.allclose(quat_slerp(q0, q0, 0.3), q0)


def test_target_trajectories_interpolation():
    u = np.zeros(3)
    tt = TargetTrajectories([0.0, 2.0], [np.r_[0, 0, 0, 0, 0, 0, 1, 0], np.r_[2, 4, 6, 0, 0, 0, 1, 0]], [u, u])
    assert np.allclose(tt.get_desired_state(1.0)[:3], [1, 2, 3])
    assert np.allclose(tt.get_desired_state(-1.0)[:3], [0, 0, 0])
    assert np.allclose(tt.get_desired_state(5.0)[:3], [2, 4, 6])
    assert np.allclose(tt.positions_at([0.5, 1.5]), [[0.5, 1, 1.5], [1.5, 3, 4.5]])
    cfg_wp = {"waypoints": [{"time": 0, "position": [-0.25, 0.5, 0.25], "orientation": [0, 0, 0, 1]}]}
    t2 = TargetTrajectories.from_config(cfg_wp, np.array([1.0, 1, 1]), np.array([0, 0, 0, 1.0]), u)
    assert np.allclose(t2.xs[0], [0.75, 1.5, 1.25, 0, 0, 0, 1, 0])


def test_workload_is_seeded_and_shaped():
    d, meta = problem_io.load_fixture("cfg2_thing_demo")
    ee = lambda x: np.zeros((x.shape[0], 3))  # noqa: E731
    a = workload.sample_batch("cfg2_thing_demo", d, meta, 16, 5, ee)
    b = workload.sample_batch("cfg2_thing_demo", d, meta, 16, 5, ee)
    assert np.array_equal(a["x0"], b["x0"]) and np.array_equal(a["target"], b["target"])
    assert a["x0"].shape == (16, 27) and a["target"].shape == (16, 21, 3) and a["body_params"].shape == (16, 1, 10)
    assert np.all(a["x0"][:, 9:] == 0)
    assert workload.algorithmic_bytes_per_solve(d) == 3744  # cfg2 3 744 B


def test_receding_horizon_shift_with_stub_engine():
    """Warm-start shift + policy evaluation of manager._RecedingHorizon."""
    from types import SimpleNamespace

    from upright_b200.manager import _RecedingHorizon

    N, nx, nu = 4, 3, 2

    class Stub:
        def __init__(self):
            self.N, self.nx, self.nu = N, nx, nu
            self.calls = []

        def set_option(self, *a):
            pass

        def solve(self, x0, target, body, X=None, U=None, warm=False, want_gains=False, **kw):
            self.calls.append(dict(warm=warm, X=None if X is None else X.copy(), U=None if U is None else U.copy()))
            B = x0.shape[0]
            Xo = np.stack([x0 + k for k in range(N + 1)], axis=1)
            Uo = np.stack([np.full((B, nu), float(k)) for k in range(N)], axis=1)
            return dict(X=Xo, U=Uo, status=np.zeros(B, np.int32), stats=np.zeros((B, 8)), K=np.zeros((B, N, nu, nx)))

    st = SimpleNamespace(sqp=SimpleNamespace(dt=0.1, use_feedback_policy=True, init_sqp_iteration=1, sqp_iteration=1),
                         mpc=SimpleNamespace(cold_start=False))
    eng = Stub()
    rh = _RecedingHorizon(eng, st, 1)
    tt = TargetTrajectories([0.0], [np.r_[1, 2, 3, 0, 0, 0, 1, 0]], [np.zeros(nu)])
    rh.reset([tt])
    rh.observe(0.0, np.zeros((1, nx)))
    rh.advance()
    assert eng.calls[0]["warm"] is False
    x, u = rh.evaluate(0.05, np.zeros((1, nx)))
    assert np.allclose(x, 0.5) and np.allclose(u, 0.5)   # linear interpolation between knots 0 and 1
    rh.observe(0.1, np.ones((1, nx)))
    rh.advance()
    assert eng.calls[1]["warm"] is True
    # previous solution shifted by exactly one knot, tail held
    assert np.allclose(eng.calls[1]["X"][0, :, 0], [1, 2, 3, 4, 4])
    assert np.allclose(eng.calls[1]["U"][0, :, 0], [1, 2, 3, 3])


def test_end_effector_box_and_next_rows():
    """EndEffectorBoxConstraint settings reach the C description; the 'next' rows are rejected loudly."""
    import copy
    d, meta = problem_io.load_fixture("cfg2_thing_demo")
    cfg = copy.deepcopy(meta["controller_config"])
    cfg["end_effector_box_constraint"] = {"enabled": True, "xyz_lower": [-1.0, -1.0, -0.05], "xyz_upper": [1.0, 1.0, 0.05]}
    desc = settings.ControllerSettings(cfg, x0=np.array(meta["x0"])).to_desc()
    assert desc.ee_box_enabled == 1
    assert list(desc.ee_box_lower) == [-1.0, -1.0, -0.05] and list(desc.ee_box_upper) == [1.0, 1.0, 0.05]
    bad = copy.deepcopy(cfg)
    bad["end_effector_box_constraint"]["xyz_lower"] = [2.0, -1.0, -0.05]
    with pytest.raises(ValueError):
        settings.ControllerSettings(bad, x0=np.array(meta["x0"])).to_desc()


def test_inertial_alignment_cost_settings():
    """InertialAlignmentCostGaussNewton settings (wrappers.py:332-345) reach the C description."""
    import copy
    from upright_b200 import geometry as geo
    d, meta = problem_io.load_fixture("cfg2_thing_demo")
    cfg = copy.deepcopy(meta["controller_config"])
    cfg["inertial_alignment"] = {"cost_enabled": True, "constraint_enabled": False, "use_angular_acceleration": False,
                                 "align_with_fixed_vector": False, "cost_weight": 2.5, "contact_plane_normal": [0, 0, 2],
                                 "com": [0, 0, 0], "alpha": 0}
    desc = settings.ControllerSettings(cfg, x0=np.array(meta["x0"])).to_desc()
    assert desc.ia_cost_enabled == 1 and desc.ia_cost_weight == 2.5 and desc.ia_constraint_enabled == 0
    S = np.array(list(desc.ia_span)).reshape(2, 3)
    assert np.allclose(S, geo.plane_span([0, 0, 1])) and np.allclose(S @ [0, 0, 1], 0) and np.allclose(S @ S.T, np.eye(2))
    cfg["inertial_alignment"].update(cost_enabled=False, constraint_enabled=True, alpha=0.2, com=[0.0, 0.0, 0.1],
                                     use_angular_acceleration=True)
    d2 = settings.ControllerSettings(cfg, x0=np.array(meta["x0"])).to_desc()
    assert (d2.ia_cost_enabled, d2.ia_constraint_enabled, d2.ia_use_angular_acceleration, d2.ia_align_with_fixed_vector) == (0, 1, 1, 0)
    assert d2.ia_alpha == 0.2 and list(d2.ia_normal) == [0.0, 0.0, 1.0] and list(d2.ia_com) == [0.0, 0.0, 0.1]


def test_dynamic_obstacle_settings():
    """`obstacles.dynamic` (wrappers.py:363-384): +9 states per obstacle, a sphere riding on each, initial state at
    rest at the first mode."""
    import copy
    d, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    cfg = copy.deepcopy(meta["controller_config"])
    cfg["obstacles"]["dynamic"] = [{"name": "ball", "radius": 0.1,
                                    "modes": [{"time": 0, "position": [3.0, 0.5, 0.8], "velocity": [-1.0, 0, 0], "acceleration": [0, 0, -9.81]}]}]
    cfg["obstacles"]["collision_pairs"] = list(cfg["obstacles"]["collision_pairs"]) + [["balanced_object_collision_link_0", "ball"]]
    s = settings.ControllerSettings(cfg)
    assert s.dims.o == 1 and s.dims.x() == 36 and s.initial_state.shape == (36,)
    assert np.allclose(s.initial_state[27:], [3.0, 0.5, 0.8, 0, 0, 0, 0, 0, 0])
    desc = s.to_desc()
    assert desc.n_dynamic_obstacles == 1 and desc.n_pairs == d.n_pairs + 1
    riding = [i for i in range(desc.n_spheres) if desc.spheres[i].link == -2]
    assert len(riding) == 1 and desc.spheres[riding[0]].radius == 0.1


@pytest.mark.skipif(not REF.exists(), reason="reference tree only in the build container")
@pytest.mark.parametrize("path", ["upright_cmd/config/ral23/experiments/sudden_obstacle/sudden_t1.0.yaml",
                                  "upright_cmd/config/ral23/experiments/projectile/projectile_head_on.yaml"])
def test_reference_obstacle_configs_load_unchanged(path):
    """The ral23 obstacle experiment families pair a wrist sphere against `ground` — the half-space the reference adds
    to every collision model (add_ground_plane, controller_interface.cpp:93-101,189).  The configurations load as
    shipped and the pair becomes a sphere / half-space row."""
    cfg = config.load_config(REF / path)["controller"]
    s = settings.ControllerSettings(cfg)
    desc = s.to_desc()
    assert desc.obstacles_enabled == 1 and desc.n_dynamic_obstacles == 1
    hs = [i for i in range(desc.n_spheres) if desc.spheres[i].shape == B.UB_SHAPE_HALFSPACE]
    assert len(hs) == 1 and desc.spheres[hs[0]].link == -1 and desc.spheres[hs[0]].radius == 0.0
    assert list(desc.spheres[hs[0]].offset) == [0.0, 0.0, 1.0]
    ground_pairs = [(desc.pairs[i].a, desc.pairs[i].b) for i in range(desc.n_pairs) if hs[0] in (desc.pairs[i].a, desc.pairs[i].b)]
    assert len(ground_pairs) == 1 and ground_pairs[0][1] == hs[0]          # the sphere first, the half-space second
    assert desc.spheres[ground_pairs[0][0]].link >= 0                        # a robot (wrist) sphere
    assert desc.n_pairs == len(cfg["obstacles"]["collision_pairs"])
    with pytest.raises(ValueError, match="unknown collision object"):
        bad = copy.deepcopy(cfg)
        bad["obstacles"]["collision_pairs"] = list(bad["obstacles"]["collision_pairs"]) + [["wrist3_collision_link_0", "no_such_thing"]]
        settings.ControllerSettings(bad).to_desc()


def test_projectile_path_constraint_settings():
    """`projectile_path_constraint` (wrappers.py:252-265; ral23/experiments/projectile/_base.yaml:81-86): the listed
    collision links become rows measured from their collision spheres to the last dynamic obstacle; the sanity
    checks of the reference (projectile_path_constraint.h:57-62) are kept."""
    import copy
    d, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    cfg = copy.deepcopy(meta["controller_config"])
    cfg["obstacles"]["dynamic"] = [{"name": "projectile1", "radius": 0.2,
                                    "modes": [{"time": 0, "position": [0, -10, 0], "velocity": [0, 0, 0], "acceleration": [0, 0, -9.81]}]}]
    cfg["projectile_path_constraint"] = {"enabled": True, "distances": [0.35, 0.2], "scale": 0.2,
                                         "collision_links": ["balanced_object_collision_link", "wrist3_collision_link"]}
    s = settings.ControllerSettings(cfg)
    desc = s.to_desc()
    assert desc.projectile_enabled == 1 and desc.n_projectile_links == 2 and desc.n_dynamic_obstacles == 1
    assert list(desc.projectile_distances)[:2] == [0.35, 0.2] and desc.projectile_scale == 0.2 and desc.projectile_active == 0.0
    slots = list(desc.projectile_spheres)[:2]
    assert all(0 <= k < desc.n_spheres and desc.spheres[k].link >= 0 for k in slots) and slots[0] != slots[1]
    assert desc.spheres[slots[0]].link == desc.nq   # the balanced object's sphere rides on the tool frame
    assert desc.n_pairs == d.n_pairs                # the projectile has no distance rows of its own (dynamic.yaml:31-36)
    bad = copy.deepcopy(cfg)
    bad["projectile_path_constraint"]["distances"] = [0.35]
    with pytest.raises(RuntimeError):
        settings.ControllerSettings(bad).to_desc()
    bad = copy.deepcopy(cfg)
    bad["obstacles"]["dynamic"] = []
    with pytest.raises(ValueError):
        settings.ControllerSettings(bad).to_desc()


def test_operating_point_initializer(tmp_path):
    """ocs2::OperatingPoints (controller_interface.cpp:380-387, wrappers.py:289-296): the trajectory file is loaded
    by the settings; with `use_operating_points` set the first solve starts from it (u_k = inputs(t_k),
    x_{k+1} = states(t_{k+1}), x_0 observed) and, on later solves, so do the knots beyond the previous horizon."""
    import copy
    from types import SimpleNamespace

    from upright_b200.manager import _RecedingHorizon
    from upright_b200.trajectory import StateInputTrajectory

    d, meta = problem_io.load_fixture("cfg2_thing_demo")
    ts = np.array([0.0, 1.0, 3.0])
    xs = np.stack([np.full(27, v) for v in (0.0, 1.0, 2.0)])
    us = np.stack([np.full(13, v) for v in (0.0, -1.0, -1.0)])
    path = tmp_path / "operating.npz"
    StateInputTrajectory(ts, xs, us).save(path)
    st = settings.ControllerSettings(copy.deepcopy(meta["controller_config"]), x0=np.array(meta["x0"]),
                                     operating_trajectory=StateInputTrajectory.load(path))
    assert st.use_operating_points is False and len(st.operating_times) == 3    # loaded, not switched on (as upstream)
    st.to_desc()
    st.use_operating_points = True
    st.to_desc()
    bad = settings.ControllerSettings(copy.deepcopy(meta["controller_config"]), x0=np.array(meta["x0"]))
    bad.use_operating_points = True
    with pytest.raises(ValueError):
        bad.to_desc()

    N, nx, nu = 20, 27, 13

    class Stub:
        def __init__(self):
            self.N, self.nx, self.nu = N, nx, nu
            self.calls = []

        def set_option(self, *a):
            pass

        def solve(self, x0, target, body, X=None, U=None, warm=False, want_gains=False, **kw):
            self.calls.append(dict(warm=warm, X=None if X is None else X.copy(), U=None if U is None else U.copy()))
            B = x0.shape[0]
            return dict(X=np.full((B, N + 1, nx), 7.0), U=np.full((B, N, nu), 7.0), status=np.zeros(B, np.int32),
                        stats=np.zeros((B, 8)), K=None)

    eng = Stub()
    rh = _RecedingHorizon(eng, st, 2)
    rh.reset([TargetTrajectories([0.0], [np.r_[1, 2, 3, 0, 0, 0, 1, 0]], [np.zeros(nu)])])
    x_obs = np.full((2, nx), 0.25)
    rh.observe(0.5, x_obs)
    rh.advance()
    c = eng.calls[0]
    assert c["warm"] is True                                     # the guess goes in as the starting iterate
    tk = 0.5 + 0.1 * np.arange(N + 1)
    want_x = np.where(tk <= 1.0, tk, 1.0 + (tk - 1.0) / 2.0)
    assert np.allclose(c["X"][:, 0], 0.25) and np.allclose(c["X"][0, 1:, 5], want_x[1:]) and np.allclose(c["X"][1], c["X"][0])
    assert np.allclose(c["U"][0, :, 3], -np.minimum(tk[:-1], 1.0))
    # next solve 0.3 s later: covered knots from the previous solution (7), the three beyond it from the initializer
    rh.observe(0.8, x_obs)
    rh.advance()
    c = eng.calls[1]
    tn = 0.8 + 0.1 * np.arange(N + 1)
    assert np.allclose(c["X"][0, :18, 5], 7.0) and np.allclose(c["X"][0, 18:, 5], 1.0 + (tn[18:] - 1.0) / 2.0)
    assert np.allclose(c["U"][0, :17, 3], 7.0) and np.allclose(c["U"][0, 17:, 3], -1.0)


def test_ballistic_obstacles_and_projectile_gate():
    """plant.BallisticObstacles = the uncontrolled obstacle of upright_sim (simulation.py:300-435): free flight under
    the mode's acceleration, reset to the next mode when its time has come, `relative` placement; plant.ProjectileGate
    = mrt_node.cpp:241-263."""
    from upright_b200.plant import BallisticObstacles, ProjectileGate
    cfg = [{"controlled": False, "radius": 0.1, "relative": True,
            "modes": [{"time": 0, "position": [0.0, -2.0, 0.3], "velocity": [0, 2.67, 3.68], "acceleration": [0, 0, -9.81]},
                      {"time": 0.5, "position": [1.414, -1.414, 0], "velocity": [-1.89, 1.89, 3.68], "acceleration": [0, 0, -9.81]}]}]
    off = np.array([[1.0, 2.0, 0.7], [0.0, 0.0, 0.5]])
    ob = BallisticObstacles(cfg, 2, offsets=off)
    assert len(ob) == 1 and ob.state().shape == (2, 9)
    assert np.allclose(ob.state()[:, :3], off + [0.0, -2.0, 0.3]) and np.allclose(ob.state()[:, 6:], [0, 0, -9.81])
    h, resets = 0.01, []
    for s in range(60):
        resets.append(ob.step(s * h, h))
    # entered the second mode at the step starting at t = 0.5, then flew 10 steps from its initial values
    assert resets.index(True) == 50 and sum(resets) == 1
    t = 10 * h
    assert np.allclose(ob.state()[:, :3], off + [1.414, -1.414, 0] + t * np.array([-1.89, 1.89, 3.68]) + 0.5 * t * t * np.array([0, 0, -9.81]))
    assert np.allclose(ob.state()[:, 3:6], np.array([-1.89, 1.89, 3.68]) + t * np.array([0, 0, -9.81]))
    with pytest.raises(NotImplementedError):
        BallisticObstacles([dict(cfg[0], controlled=True)], 1)
    g = ProjectileGate()
    zs = [0.5, 0.9, 1.01, 1.5, 0.8, 0.21, 0.19, 1.5]
    assert [g.update(z) for z in zs] == [0, 0, 1, 1, 1, 1, 0, 0] and g.observing   # no second flight


def test_host_rollout_with_projectile_on_oracle_engine():
    """BatchedControllerManager.rollout_host — the mpc_sim.py loop with the simulated projectile feeding the obstacle
    columns and the in-flight gate raising the flag s — driven here by the CPU oracle standing in for the engine:
    the tray keeps a larger distance from the ball than with the constraint disabled."""
    import oracle
    from _util import OracleEngine, projectile_rollout
    dist = {}
    for enabled in (True, False):
        out, desc, info = projectile_rollout(OracleEngine, enabled)
        eng = info["engine"]
        assert out["xs"].shape == (1, 70, 36) and out["n_replans"] == 14 and np.isfinite(out["xs"]).all()
        # the obstacle columns are the simulated flight, the flag follows the height of the ball
        assert np.allclose(out["xs"][0, :, 27:30], info["flight"])
        z = out["xs"][0, :, 29]
        assert out["flags"][0] == 1.0 and out["flags"][-1] == 0.0 and (out["flags"][z > 1.0] == 1.0).all()
        assert eng.flags[0] == 1.0 and eng.flags[-1] == 0.0
        cen = np.array([oracle.fk(desc, out["xs"][0, k])["spheres"][info["tray"]] for k in range(70)])
        dist[enabled] = np.linalg.norm(cen - out["xs"][0, :, 27:30], axis=1).min()
    assert dist[True] > dist[False] + 0.02, dist   # measured: 0.42 m against 0.20 m


def test_fp32_conditioning_estimate():
    """The estimate that makes BatchedMPC warn: fine for the shipped configurations, beyond fp32 for cfg4's 20 g
    object once its object-dynamics rows are softened (measured on the B200: NaN status for every instance)."""
    import copy
    from upright_b200.problem_io import fp32_conditioning_estimate
    demo, _ = problem_io.load_fixture("cfg2_thing_demo")
    obst, _ = problem_io.load_fixture("cfg4_thing_obstacles2")
    assert 1e4 < fp32_conditioning_estimate(demo) < 1e7
    assert fp32_conditioning_estimate(obst) == 1.0            # slacks off: normalised hard rows
    soft = copy.deepcopy(obst)
    soft.slacks.enabled = 1
    assert fp32_conditioning_estimate(soft) > 1e8
    heavy = np.tile(np.array(list(demo.body_params[0])), (3, 1, 1))
    assert fp32_conditioning_estimate(soft, heavy) == pytest.approx(fp32_conditioning_estimate(demo))


def _oracle_closed_loop(cfg, x0, goal, duration, replan=0.05):
    """mpc_sim.py loop on the CPU oracle (see test_closed_loop_solves_the_waiters_problem_on_oracle_engine)."""
    from _util import OracleEngine
    from upright_b200.manager import BatchedControllerManager, _RecedingHorizon
    st = settings.ControllerSettings(cfg, x0=x0[0])
    d = st.to_desc()
    mgr = object.__new__(BatchedControllerManager)
    mgr.settings, mgr.desc, mgr.engine, mgr.B = st, d, OracleEngine(d), 1
    mgr.core = _RecedingHorizon(mgr.engine, st, 1)
    mgr.core.reset([TargetTrajectories([0.0], [np.r_[goal, 0, 0, 0, 1, 0]], [np.zeros(st.dims.u())])])
    mgr.core.body_params = None
    mgr.timestep, mgr.last_planning_time = replan, -np.inf
    mgr.replanning_times, mgr.replanning_durations = [], []
    return mgr.rollout_host(x0, duration, 0.01)["xs"][0]


@pytest.mark.parametrize("name", ["cfg1_ur10_demo", "cfg2_thing_demo", "cfg5_thing_robust8"])
def test_closed_loop_solves_the_waiters_problem_on_oracle_engine(name):
    """End to end on the CPU: 5 s of the mpc_sim.py loop (replan every 50 ms, warm-start shift, feedback policy) with
    the oracle as the engine, for the frictionless configurations (fixed-base arm, mobile manipulator, the robust set
    of eight CoM-vertex bodies).  The tray reaches the waypoint of thing_demo.yaml and comes to rest, and at every
    simulated state there are NON-NEGATIVE normal forces that explain the motion of every body (non-negative least
    squares on the object-dynamics rows) — while the same move planned without the balancing constraints needs
    forces that do not exist."""
    import copy

    import oracle
    from scipy.optimize import nnls

    desc, meta = problem_io.load_fixture(name)
    x0 = np.array(meta["x0"], dtype=float)[None]
    goal = np.array(meta["r_ee0"]) + [-0.25, 0.5, 0.25]            # thing_demo.yaml:55

    def missing_force(xs):
        worst = 0.0
        for x in xs[::5]:
            lin = oracle.linearize(desc, x, np.zeros(desc.nu))
            worst = max(worst, nnls(lin["Df"], -lin["g"])[1])
        return worst

    xs = _oracle_closed_loop(meta["controller_config"], x0, goal, 5.0)
    assert np.linalg.norm(oracle.fk(desc, xs[-1])["r"] - goal) < 1e-2
    assert np.abs(xs[-1, desc.nq:]).max() < 2e-2                        # at rest
    at_rest = np.abs(oracle.linearize(desc, xs[0], np.zeros(desc.nu))["g"]).max()   # gravity on the scaled rows
    assert missing_force(xs) < 0.05 * at_rest      # measured 0.014 / 0.013 / 0.051 against 4.0 / 4.0 / 1.42
    free = copy.deepcopy(meta["controller_config"])
    free["balancing"]["enabled"] = False
    xs_free = _oracle_closed_loop(free, x0, goal, 5.0)
    assert np.linalg.norm(oracle.fk(desc, xs_free[-1])["r"] - goal) < 5e-3
    assert missing_force(xs_free) > 0.1 * at_rest  # measured 0.96 / 1.07 / 1.13


def test_closed_loop_keeps_friction_cones_on_oracle_engine():
    """cfg3 (three stacked bodies, 16 friction contacts): along 4 s of closed loop there are contact forces inside
    the friction pyramids that carry all three bodies at every simulated state (a linear programme on the
    object-dynamics rows and the pyramid rows); planned without the balancing constraints there are none."""
    import copy

    import oracle
    from scipy.optimize import linprog
    desc, meta = problem_io.load_fixture("cfg3_thing_box_arch")
    x0 = np.array(meta["x0"], dtype=float)[None]
    goal = np.array(meta["r_ee0"]) + [-0.25, 0.5, 0.25]
    nfc = desc.nf * desc.nc

    def cone_gap(xs):
        worst = 0.0
        for x in xs[::20]:
            lin = oracle.linearize(desc, x, np.zeros(desc.nq + nfc))
            Df, g, F = lin["Df"], lin["g"], lin["Ffric"]
            ne, one = Df.shape[0], np.ones((Df.shape[0], 1))
            res = linprog(np.r_[np.zeros(nfc), 1.0],                       # min t: |Df f + g| <= t, F f >= 0
                          A_ub=np.block([[Df, -one], [-Df, -one], [-F, np.zeros((F.shape[0], 1))]]),
                          b_ub=np.r_[-g, g, np.zeros(F.shape[0])], bounds=[(-100, 100)] * nfc + [(0, None)])
            assert res.status == 0
            worst = max(worst, res.x[-1])
        return worst

    xs = _oracle_closed_loop(meta["controller_config"], x0, goal, 4.0)
    assert np.linalg.norm(oracle.fk(desc, xs[-1])["r"] - goal) < 1e-2
    assert cone_gap(xs) < 1e-3                                             # measured 0
    free = copy.deepcopy(meta["controller_config"])
    free["balancing"]["enabled"] = False
    assert cone_gap(_oracle_closed_loop(free, x0, goal, 4.0)) > 0.05       # measured 0.11


def test_closed_loop_avoids_obstacles_on_oracle_engine():
    """cfg4 (hard sphere-distance rows): sent to a point behind an obstacle the closed loop keeps every collision
    pair at its minimum distance and stops in front of it (a 2 s horizon does not plan around), where the same loop
    with obstacles disabled drives through."""
    import copy

    import oracle
    desc, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    x0 = np.array(meta["x0"], dtype=float)[None]
    goal = np.array(meta["r_ee0"]) + [1.9, -0.25, 0.0]
    margin = lambda xs: min(oracle.linearize(desc, x, np.zeros(desc.nu))["hobs"].min() for x in xs[::5])  # noqa: E731
    xs = _oracle_closed_loop(meta["controller_config"], x0, goal, 6.0)
    d0 = np.linalg.norm(oracle.fk(desc, xs[0])["r"] - goal)
    # Near the obstacle the hard distance rows make the QPs end at the iteration cap, and the closed loop is then
    # sensitive to rounding: the same sources compiled with / without FMA contraction, or with unrelated code added to
    # the translation unit, stop at margins of +0.084, +0.011 or -0.0010 m.  What holds in all of them: the pairs keep
    # their minimum distance to within 2 mm (against -0.48 m with the rows disabled).
    assert margin(xs) > -2e-3
    assert np.linalg.norm(oracle.fk(desc, xs[-1])["r"] - goal) < 0.6 * d0  # it does approach
    free = copy.deepcopy(meta["controller_config"])
    free["obstacles"]["enabled"] = False
    xs_free = _oracle_closed_loop(free, x0, goal, 6.0)
    assert margin(xs_free) < -0.2                                          # measured -0.48
    assert np.linalg.norm(oracle.fk(desc, xs_free[-1])["r"] - goal) < 1e-2


def test_closed_loop_inertial_alignment_on_oracle_engine():
    """Inertial alignment as the alternative to the balancing constraints (inertial_alignment.h:68-204): with the
    five constraint rows the tray normal stays inside the alpha-pyramid around C_we'(a - g) along the whole closed
    loop, with the Gauss-Newton cost alone it leans that way, with neither the rows are violated grossly."""
    import copy

    import oracle
    desc, meta = problem_io.load_fixture("cfg2_thing_demo")
    x0 = np.array(meta["x0"], dtype=float)[None]
    goal = np.array(meta["r_ee0"]) + [-0.25, 0.5, 0.25]

    def cfg(constraint, cost):
        c = copy.deepcopy(meta["controller_config"])
        c["balancing"]["enabled"] = False
        c["inertial_alignment"] = {"cost_enabled": cost, "constraint_enabled": constraint, "use_angular_acceleration": False,
                                   "align_with_fixed_vector": False, "cost_weight": 10, "contact_plane_normal": [0, 0, 1],
                                   "com": [0, 0, 0], "alpha": 0.05}
        return c

    probe = settings.ControllerSettings(cfg(True, False), x0=x0[0]).to_desc()

    def worst_row(xs):
        w = np.inf
        for x in xs[::5]:   # the five rows at state x = the constants of the QP rows of a trajectory resting at x
            q = oracle.qp_dump(probe, np.tile(goal, (probe.N + 1, 1)), np.tile(x, (probe.N + 1, 1)), np.zeros((probe.N, 9)))[3]
            w = min(w, q["c"][-5:].min())
        return w

    res = {}
    for label, c in (("constraint", cfg(True, False)), ("cost", cfg(False, True)), ("neither", cfg(False, False))):
        xs = _oracle_closed_loop(c, x0, goal, 5.0)
        assert np.linalg.norm(oracle.fk(probe, xs[-1])["r"] - goal) < 5e-3
        res[label] = worst_row(xs)
    assert res["constraint"] > -0.02                     # measured -0.004 (soft rows)
    assert res["neither"] < -1.0                         # measured -3.0
    assert res["neither"] < res["cost"] < res["constraint"]   # measured -0.59
