def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def projectile_problem(active=1.0, light_object=False):
    """cfg4 plus one dynamic obstacle (the projectile, no collision pairs of its own, as in
    upright_cmd/config/obstacles/dynamic.yaml:18-36) and projectile-path rows for the tray sphere and one arm sphere
    (ral23/experiments/projectile/_base.yaml:81-86), softened like the other inequalities (thing_demo.yaml:31-33).
    Returns (desc, meta, tray sphere slot)."""
    import copy
    from upright_b200 import problem_io
    desc, meta = problem_io.load_fixture("cfg4_thing_obstacles2")
    d = copy.deepcopy(desc)
    # the balanced object of thing_demo.yaml (cfg2): cfg4's 20 g object scales the soft object-dynamics rows by
    # 1/m = 50, which costs the fp32 kernels four digits of conditioning against the force weight
    # (`light_object=True` keeps it: the documented fp32 breakdown case)
    demo, _ = problem_io.load_fixture("cfg2_thing_demo")
    assert (demo.nb, demo.nc, demo.nf) == (d.nb, d.nc, d.nf)
    if not light_object:
        for i in range(len(d.body_params[0])):
            d.body_params[0][i] = demo.body_params[0][i]
        for i in range(d.nc):
            d.contacts[i] = demo.contacts[i]
    d.n_dynamic_obstacles = 1
    robot = [i for i in range(d.n_spheres) if d.spheres[i].link >= 0]
    tray = max(robot, key=lambda i: d.spheres[i].link)
    d.projectile_enabled, d.n_projectile_links = 1, 2
    d.projectile_spheres[0], d.projectile_spheres[1] = tray, robot[0]
    d.projectile_distances[0], d.projectile_distances[1] = 0.35, 0.2
    d.projectile_scale, d.projectile_active = 0.2, active
    d.slacks.enabled, d.slacks.poly_ineq = 1, 1
    return d, meta, tray


def ballistic_prediction(xo, N, dt):
    """Constant-acceleration rollout [N+1, 9] of obstacle states xo = [p, v, a] (system_dynamics.h:28-38)."""
    import numpy as np
    xo = np.asarray(xo, dtype=float)
    t = dt * np.arange(N + 1)[:, None]
    return np.hstack((xo[None, :3] + t * xo[None, 3:6] + 0.5 * t * t * xo[None, 6:], xo[None, 3:6] + t * xo[None, 6:],
                      np.tile(xo[None, 6:], (N + 1, 1))))


def projectile_throws(centre, gravity=(0.0, 0.0, -9.81)):
    """Obstacle states [p, v, a] of balls released above the tray sphere `centre` and descending past it."""
    import numpy as np
    g = np.asarray(gravity, dtype=float)
    out = []
    for T, off, start in ((0.45, [0.2, 0, 0.0], [0.3, -2.0, 1.0]), (0.45, [0.0, 0.0, 0.25], [0.3, -2.0, 1.0]),
                          (0.4, [0.5, 0, 0.1], [-0.5, -1.5, 0.9]), (0.3, [-0.1, 0, 0.15], [1.0, 1.0, 0.5])):
        p0 = centre + np.array(start)
        v0 = (centre + np.array(off) - p0) / T - 0.5 * g * T
        out.append(np.concatenate((p0, v0, g)))
    return np.array(out)


class OracleEngine:
    """The CPU oracle behind the engine interface the receding-horizon core uses (host-logic tests only)."""

    def __init__(self, desc):
        import oracle
        self.desc = desc
        dm = oracle.dims(desc)
        self.N, self.nx, self.nu, self.nx_robot = dm["N"], dm["nx"], dm["nu"], 3 * desc.nq
        self.flags = []

    def set_option(self, key, value):
        if key == "projectile_active":
            self.desc.projectile_active = float(value)
        elif key == "sqp_iteration":
            self.desc.sqp_iteration = int(value)

    def solve(self, x0, target, body, X=None, U=None, warm=False, want_gains=False, **kw):
        import oracle
        self.flags.append(self.desc.projectile_active)
        return oracle.solve_batch(self.desc, x0, target, body, X=X, U=U, warm=warm, want_gains=want_gains)


def projectile_rollout(make_engine, enabled, duration=0.7, sim_dt=0.01):
    """0.7 s of the mpc_sim.py loop on the projectile problem: a ball released 1 m above and 2 m beside the tray
    passes 0.2 m from the tray sphere after 0.45 s.  `make_engine(desc)` supplies the solver.  Returns the rollout,
    the problem description and {engine, tray, flight [n, 3]}."""
    import copy
    import numpy as np
    import oracle
    from upright_b200 import settings
    from upright_b200.manager import BatchedControllerManager, _RecedingHorizon
    from upright_b200.plant import BallisticObstacles, ProjectileGate
    from upright_b200.settings import TargetTrajectories
    base, meta, tray = projectile_problem(active=0.0)
    base.projectile_scale = 2.0
    desc = copy.deepcopy(base)
    desc.projectile_enabled = int(enabled)
    x0 = np.array(meta["x0"], dtype=float)[None]
    c = oracle.fk(base, np.concatenate((x0[0], np.zeros(9))))["spheres"][tray]
    g = np.array([0.0, 0.0, -9.81])
    T, start, aim = 0.45, np.array([0.3, -2.0, 1.0]), np.array([0.2, 0.0, 0.0])
    v0 = (aim - start) / T - 0.5 * g * T
    sim = [{"controlled": False, "radius": 0.1, "relative": True,
            "modes": [{"time": 0, "position": list(start), "velocity": list(v0), "acceleration": list(g)}]}]
    st = settings.ControllerSettings(meta["controller_config"], x0=x0[0])
    st.projectile_path_constraint_enabled = True     # read by the receding-horizon core (flag forwarding)
    st.sqp.use_feedback_policy = False
    mgr = object.__new__(BatchedControllerManager)       # the real constructor builds its own description
    mgr.settings, mgr.desc, mgr.engine, mgr.B = st, desc, make_engine(desc), 1
    mgr.core = _RecedingHorizon(mgr.engine, st, 1)
    mgr.core.reset([TargetTrajectories([0.0], [np.r_[meta["r_ee0"], 0, 0, 0, 1, 0]], [np.zeros(13)])])
    mgr.core.body_params = None
    mgr.timestep, mgr.last_planning_time = 0.05, -np.inf
    mgr.replanning_times, mgr.replanning_durations = [], []
    out = mgr.rollout_host(x0, duration, sim_dt, obstacles=BallisticObstacles(sim, 1, offsets=c[None]), gate=ProjectileGate())
    tt = sim_dt * np.arange(int(round(duration / sim_dt)))[:, None]
    return out, desc, dict(engine=mgr.engine, tray=tray, flight=c + start + tt * v0 + 0.5 * tt * tt * g)
