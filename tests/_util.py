def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def projectile_problem(active=1.0):
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
    demo, _ = problem_io.load_fixture("cfg2_thing_demo")
    assert (demo.nb, demo.nc, demo.nf) == (d.nb, d.nc, d.nf)
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
