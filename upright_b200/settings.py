"""`ControllerSettings` / `TargetTrajectories` with the attribute surface of the
reference's pybind structs, filled from the same YAML dictionaries.

Mirrors `upright_control/src/upright_control/wrappers.py:14-399` (field names,
defaults, assertions) and the struct tree of
`upright_control/include/upright_control/controller_settings.h:47-119`.  The
URDF step (`parse_and_compile_urdf`, wrappers.py:279-281) is replaced by the
kinematic-chain fixture in `robot.py`; everything else is parsed identically.
`to_desc()` flattens the settings into the C-ABI `ub_problem_desc_t`.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from types import SimpleNamespace

import numpy as np

from . import bindings as B
from . import config as cfgmod
from . import geometry as geo
from . import objects, robot


class RobotDimensions(SimpleNamespace):
    def __init__(self):
        super().__init__(q=0, v=0, x=0, u=0)


class OptimizationDimensions:
    """upright_control/include/upright_control/dimensions.h:18-46."""

    def __init__(self):
        self.robot = RobotDimensions()
        self.o = 0
        self.c = 0
        self.nf = 3

    def q(self):
        return self.robot.q + 3 * self.o

    def v(self):
        return self.robot.v + 3 * self.o

    def x(self):
        return self.robot.x + 9 * self.o

    def f(self):
        return self.nf * self.c

    def u(self):
        return self.robot.u + self.f()


class TargetTrajectories:
    """ocs2::TargetTrajectories as wrapped by wrappers.py:14-76: each target
    state is [r(3), quat xyzw(4), s(1)]."""

    def __init__(self, ts, xs, us):
        self.ts = [float(t) for t in ts]
        self.xs = [np.array(x, dtype=float) for x in xs]
        self.us = [np.array(u, dtype=float) for u in us]

    @classmethod
    def from_config(cls, config, r_ew_w, Q_we, u):
        ts, xs, us = [], [], []
        for wp in config["waypoints"]:
            r = np.asarray(r_ew_w, dtype=float) + np.asarray(wp["position"], dtype=float)
            Q = geo.quat_multiply(Q_we, np.asarray(wp["orientation"], dtype=float))
            ts.append(wp["time"])
            xs.append(np.concatenate((r, Q, [0.0])))
            us.append(np.copy(u))
        return cls(ts, xs, us)

    def get_desired_state(self, t):
        """Linear interpolation in time, clamped (ocs2 LinearInterpolation)."""
        if len(self.xs) == 1 or t <= self.ts[0]:
            return self.xs[0].copy()
        if t >= self.ts[-1]:
            return self.xs[-1].copy()
        i = int(np.searchsorted(self.ts, t, side="right")) - 1
        a = (self.ts[i + 1] - t) / (self.ts[i + 1] - self.ts[i])
        return a * self.xs[i] + (1 - a) * self.xs[i + 1]

    def get_desired_input(self, t):
        return self.us[0].copy()

    def get_desired_pose(self, t):
        x = self.get_desired_state(t)
        return x[:3], x[3:7]

    def poses(self):
        for x in self.xs:
            yield x[:3], x[3:7]

    def positions_at(self, times):
        """Desired EE position at each time (reference_trajectory.h:18-47,
        position part)."""
        return np.array([self.get_desired_state(t)[:3] for t in times])

    def poses_at(self, times):
        """[r(3), quat xyzw(4)] at each time: interpolate_end_effector_pose (reference_trajectory.h:18-47) — position
        linear, orientation `q_lhs.slerp(1 - alpha, q_rhs)` with Eigen's slerp (shortest arc; linear blend when the
        quaternions are nearly parallel)."""
        out = np.empty((len(times), 7))
        for n, t in enumerate(times):
            if len(self.xs) == 1 or t <= self.ts[0]:
                out[n] = self.xs[0][:7]
                continue
            if t >= self.ts[-1]:
                out[n] = self.xs[-1][:7]
                continue
            i = int(np.searchsorted(self.ts, t, side="right")) - 1
            a = (self.ts[i + 1] - t) / (self.ts[i + 1] - self.ts[i])
            out[n, :3] = a * self.xs[i][:3] + (1 - a) * self.xs[i + 1][:3]
            out[n, 3:] = quat_slerp(self.xs[i][3:7], self.xs[i + 1][3:7], 1 - a)
        return out


def quat_slerp(q0, q1, t):
    """Eigen::Quaternion::slerp(t, other) for [x y z w] quaternions."""
    q0, q1 = np.asarray(q0, dtype=float), np.asarray(q1, dtype=float)
    d = float(q0 @ q1)
    ad = abs(d)
    if ad >= 1.0 - np.finfo(float).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        s0, s1 = np.sin((1.0 - t) * th) / np.sin(th), np.sin(t * th) / np.sin(th)
    if d < 0:
        s1 = -s1
    return s0 * q0 + s1 * q1


def _read_obstacle_xacro(path):
    """Sphere obstacles from an `obstacle_link` xacro scene such as
    upright_assets/thing/xacro/obstacles/simple.urdf.xacro:41-102."""
    out = {}
    for el in ET.parse(path).getroot().iter():
        if el.tag.endswith("obstacle_link") and "name" in el.attrib:
            origin = el.find("origin")
            sph = el.find("geometry/sphere")
            if origin is None or sph is None:
                continue
            xyz = [float(v) for v in origin.attrib.get("xyz", "0 0 0").split()]
            out[el.attrib["name"]] = (np.array(xyz), float(sph.attrib["radius"]))
    return out


def _resolve_find(path_expr):
    # "$(find pkg)/rest" -> file path via the package roots
    assert path_expr.startswith("$(find ")
    pkg, rest = path_expr[len("$(find "):].split(")", 1)
    return cfgmod.resolve_package_path({"package": pkg, "path": rest.lstrip("/")})


class ControllerSettings:
    def __init__(self, config, x0=None, operating_trajectory=None):
        pn, pa = cfgmod.parse_number, cfgmod.parse_array
        self.config = config
        self.mpc = SimpleNamespace(
            time_horizon=pn(config["mpc"]["time_horizon"]),
            debug_print=config["mpc"]["debug_print"],
            cold_start=config["mpc"]["cold_start"],
        )
        ro = config["rollout"]
        self.rollout = SimpleNamespace(
            abs_tol_ode=pn(ro["abs_tol_ode"]), rel_tol_ode=pn(ro["rel_tol_ode"]),
            timestep=pn(ro["timestep"]),
            max_num_steps_per_second=pn(ro["max_num_steps_per_second"], dtype=int),
            check_numerical_stability=ro["check_numerical_stability"],
        )
        sq = config["sqp"]
        sl = sq["hpipm"]["slacks"]
        slacks = SimpleNamespace(
            enabled=sl["enabled"], input_box=sl.get("input_box", True),
            state_box=sl.get("state_box", True), poly_ineq=sl.get("poly_ineq", True),
            upper_L2_penalty=sl.get("upper_L2_penalty", 100),
            lower_L2_penalty=sl.get("lower_L2_penalty", 100),
            upper_L1_penalty=sl.get("upper_L1_penalty", 0),
            lower_L1_penalty=sl.get("lower_L1_penalty", 0),
            upper_low_bound=sl.get("upper_low_bound", 0),
            lower_low_bound=sl.get("lower_low_bound", 0),
        )
        self.sqp = SimpleNamespace(
            dt=pn(sq["dt"]), sqp_iteration=sq["sqp_iteration"],
            init_sqp_iteration=sq["init_sqp_iteration"], delta_tol=pn(sq["delta_tol"]),
            cost_tol=pn(sq["cost_tol"]), use_feedback_policy=sq["use_feedback_policy"],
            project_state_input_equality_constraints=sq["project_state_input_equality_constraints"],
            print_solver_status=sq["print_solver_status"],
            print_solver_statistics=sq["print_solver_statistics"],
            print_line_search=sq["print_line_search"],
            hpipm=SimpleNamespace(warm_start=sq["hpipm"]["warm_start"],
                                  iter_max=sq["hpipm"]["iter_max"], slacks=slacks),
        )
        self.end_effector_link_name = config["robot"]["tool_link_name"]
        self.robot_base_type = config["robot"]["base_type"]
        est = config.get("estimation", {})
        self.estimation = SimpleNamespace(**est)
        self.tracking = SimpleNamespace(**config.get("tracking", {}))
        self.gravity = np.array(config["gravity"], dtype=float)
        self.recompile_libraries = config.get("recompile_libraries", True)
        self.debug = config.get("debug", False)

        self.dims = OptimizationDimensions()
        d = config["robot"]["dims"]
        self.dims.robot.q, self.dims.robot.v = d["q"], d["v"]
        self.dims.robot.x, self.dims.robot.u = d["x"], d["u"]
        rq, rx, ru = self.dims.robot.q, self.dims.robot.x, self.dims.robot.u

        w = config["weights"]
        self.input_weight = cfgmod.parse_diag_matrix_dict(w["input"])
        self.state_weight = cfgmod.parse_diag_matrix_dict(w["state"])
        self.end_effector_weight = cfgmod.parse_diag_matrix_dict(w["end_effector"])
        assert self.input_weight.shape == (ru, ru)
        assert self.state_weight.shape == (rx, rx)
        assert self.end_effector_weight.shape == (6, 6)

        lim = config["limits"]
        self.input_limit_lower = pa(lim["input"]["lower"])
        self.input_limit_upper = pa(lim["input"]["upper"])
        self.state_limit_lower = pa(lim["state"]["lower"])
        self.state_limit_upper = pa(lim["state"]["upper"])
        assert self.input_limit_lower.shape == (ru,) and self.input_limit_upper.shape == (ru,)
        assert self.state_limit_lower.shape == (rx,) and self.state_limit_upper.shape == (rx,)

        ebc = config.get("end_effector_box_constraint", {"enabled": False})
        self.end_effector_box_constraint_enabled = ebc["enabled"]
        self.xyz_lower = pa(ebc.get("xyz_lower", [-1, -1, -1]))
        self.xyz_upper = pa(ebc.get("xyz_upper", [1, 1, 1]))
        ppc = config.get("projectile_path_constraint", {"enabled": False})
        self.projectile_path_constraint_enabled = ppc["enabled"]
        # wrappers.py:252-265
        self.projectile_path_distances = np.array(ppc.get("distances", []), dtype=float)
        self.projectile_path_scale = float(ppc.get("scale", 1.0))
        self.projectile_path_collision_links = list(ppc.get("collision_links", []))
        if self.projectile_path_constraint_enabled:
            assert (self.projectile_path_distances >= 0).all()
            assert self.projectile_path_scale >= 0

        self.locked_joints = {
            k: pn(v) for k, v in config["robot"].get("locked_joints", {}).items()
        }
        self.base_pose = np.array(config["robot"].get("base_pose", [0, 0, 0]), dtype=float)
        assert self.base_pose.shape == (3,)
        # The un-vendored URDF is replaced by the in-repo chain fixture.
        self.robot_urdf_path = "<upright_b200.robot fixture>"
        self.lib_folder = "/tmp/ocs2"
        # Operating points (wrappers.py:289-296; controller_settings.h:90-97; initializer chosen at
        # controller_interface.cpp:380-387).  As in the reference, loading the trajectory does not switch the
        # initializer on: `use_operating_points` stays False until the caller sets it.
        self.use_operating_points = False
        self.operating_times, self.operating_states, self.operating_inputs = [], [], []
        op = config.get("operating_points", {"enabled": False})
        if op.get("enabled", False) or operating_trajectory is not None:
            if operating_trajectory is None:
                from .trajectory import StateInputTrajectory
                operating_trajectory = StateInputTrajectory.load(cfgmod.resolve_package_path(op))
            for i in range(len(operating_trajectory)):
                self.operating_times.append(float(operating_trajectory.ts[i]))
                self.operating_states.append(np.array(operating_trajectory.xs[i], dtype=float))
                self.operating_inputs.append(np.array(operating_trajectory.us[i], dtype=float))

        bal = config["balancing"]
        self.balancing_settings = SimpleNamespace(
            enabled=bal["enabled"], arrangement_name=bal["arrangement"],
            force_weight=bal["force_weight"], bodies={}, contacts=[],
        )
        bodies, contacts = objects.parse_control_objects(config)
        self.balancing_settings.bodies = bodies
        self.balancing_settings.contacts = contacts
        if self.balancing_settings.enabled:
            self.dims.c = len(contacts)
            self.dims.nf = 1 if bal["frictionless"] else 3
        else:
            self.dims.c = 0
            self.dims.nf = 0

        ia = config.get("inertial_alignment", {})
        self.inertial_alignment_settings = SimpleNamespace(
            cost_enabled=ia.get("cost_enabled", False),
            constraint_enabled=ia.get("constraint_enabled", False),
        )
        if ia.get("cost_enabled") or ia.get("constraint_enabled"):   # wrappers.py:332-345
            ias = self.inertial_alignment_settings
            ias.use_angular_acceleration = ia.get("use_angular_acceleration", False)
            ias.align_with_fixed_vector = ia.get("align_with_fixed_vector", False)
            ias.cost_weight = float(ia.get("cost_weight", 1.0))
            normal = np.array(ia.get("contact_plane_normal", [0, 0, 1]), dtype=float)
            ias.contact_plane_normal = normal / np.linalg.norm(normal)
            ias.contact_plane_span = geo.plane_span(ias.contact_plane_normal)
            ias.com = np.array(ia.get("com", [0, 0, 0]), dtype=float)
            ias.alpha = ia.get("alpha", 0)

        obs = config.get("obstacles", {"enabled": False})
        self.obstacle_settings = SimpleNamespace(
            enabled=obs["enabled"], collision_link_pairs=[],
            minimum_distance=obs.get("minimum_distance", 0.1),
            obstacle_urdf_path="", dynamic_obstacles=[], static_spheres={},
        )
        if self.obstacle_settings.enabled:
            for pair in obs.get("collision_pairs") or []:
                self.obstacle_settings.collision_link_pairs.append(tuple(pair))
            if "spheres" in obs:  # this repo's explicit form
                for s in obs["spheres"]:
                    self.obstacle_settings.static_spheres[s["name"]] = (
                        np.array(s["position"], dtype=float), float(s["radius"]))
            elif "urdf" in obs:
                for inc in obs["urdf"]["includes"]:
                    path = _resolve_find(inc)
                    self.obstacle_settings.obstacle_urdf_path = str(path)
                    self.obstacle_settings.static_spheres.update(_read_obstacle_xacro(path))
            # dynamic obstacles: +9 states each (wrappers.py:363-384)
            for od in obs.get("dynamic") or []:
                modes = [SimpleNamespace(time=float(m["time"]), position=np.array(m["position"], dtype=float),
                                         velocity=np.array(m["velocity"], dtype=float),
                                         acceleration=np.array(m["acceleration"], dtype=float)) for m in od["modes"]]
                self.obstacle_settings.dynamic_obstacles.append(
                    SimpleNamespace(name=od["name"], radius=float(od["radius"]), modes=modes))
            self.dims.o = len(self.obstacle_settings.dynamic_obstacles)

        if x0 is None:
            x0_robot = pa(config["robot"]["x0"])
            assert x0_robot.shape == (rx,)
            # the obstacles start at rest at their first mode (wrappers.py:378-384)
            x0_obs = [np.concatenate((o.modes[0].position, np.zeros(6))) for o in self.obstacle_settings.dynamic_obstacles]
            self.initial_state = np.concatenate([x0_robot] + x0_obs)
        else:
            self.initial_state = np.array(x0, dtype=float)
        assert self.initial_state.shape == (self.dims.x(),)
        self.xd = np.array(config.get("desired_state", np.zeros_like(self.initial_state)), dtype=float)

        # kinematic chain replacing the URDF + Pinocchio model
        self.chain = robot.build_chain(
            "fixed" if self.robot_base_type == "fixed" else "omnidirectional", self.base_pose)
        assert self.chain.nq == rq, f"chain has {self.chain.nq} joints, config says {rq}"

    @classmethod
    def from_config_file(cls, path):
        return cls(cfgmod.load_config(path)["controller"])

    # ------------------------------------------------------------------ C ABI
    def body_names(self):
        """std::map iteration order of bodies (contact_constraints.h:180)."""
        return sorted(self.balancing_settings.bodies.keys())

    def to_desc(self) -> B.ProblemDesc:
        d = B.ProblemDesc()
        dims = self.dims
        nq = dims.robot.q
        d.nq = nq
        names = self.body_names() if self.balancing_settings.enabled else []
        d.nb = len(names)
        d.nc = dims.c
        d.nf = dims.nf if dims.c > 0 else 1
        if d.nb > B.UB_MAX_BODIES or d.nc > B.UB_MAX_CONTACTS:
            raise ValueError("arrangement exceeds UB_MAX_BODIES / UB_MAX_CONTACTS")
        d.dt = self.sqp.dt
        d.N = int(round(self.mpc.time_horizon / self.sqp.dt))
        d.sqp_iteration = int(self.sqp.sqp_iteration)
        d.qp_iter_max = int(self.sqp.hpipm.iter_max)
        d.balancing_enabled = int(bool(self.balancing_settings.enabled) and d.nb > 0)
        d.obstacles_enabled = int(bool(self.obstacle_settings.enabled))

        for i, j in enumerate(self.chain.joints):
            d.joints[i].type = j.type
            d.joints[i].R[:] = j.R.ravel()
            d.joints[i].p[:] = j.p
            d.joints[i].axis[:] = j.axis
        d.tool_R[:] = self.chain.tool_R.ravel()
        d.tool_p[:] = self.chain.tool_p
        d.gravity[:] = self.gravity

        # features of the reference that are 'next' rows are rejected loudly, never ignored
        ia = self.inertial_alignment_settings
        # InertialAlignmentCostGaussNewton / InertialAlignmentConstraint (controller_interface.cpp:296-315)
        d.ia_cost_enabled = int(bool(ia.cost_enabled))
        d.ia_constraint_enabled = int(bool(ia.constraint_enabled))
        if d.ia_cost_enabled or d.ia_constraint_enabled:
            d.ia_cost_weight = float(ia.cost_weight)
            d.ia_span[:] = np.asarray(ia.contact_plane_span, dtype=float).reshape(6)
            d.ia_normal[:] = ia.contact_plane_normal
            d.ia_com[:] = ia.com
            d.ia_alpha = float(ia.alpha)
            d.ia_use_angular_acceleration = int(bool(ia.use_angular_acceleration))
            d.ia_align_with_fixed_vector = int(bool(ia.align_with_fixed_vector))
        if self.use_operating_points:   # host-side initial guess (manager.py); the device needs nothing
            if not self.operating_times:
                raise ValueError("use_operating_points is set but no operating trajectory was loaded")
            if any(x.shape != (self.dims.x(),) for x in self.operating_states) or \
                    any(u.shape != (self.dims.u(),) for u in self.operating_inputs):
                raise ValueError("operating states / inputs must have the problem's state / input dimension")
        # EndEffectorBoxConstraint (end_effector_box_constraint.h; wrappers.py:240-250)
        d.ee_box_enabled = int(bool(self.end_effector_box_constraint_enabled))
        if d.ee_box_enabled:
            lo, hi = np.asarray(self.xyz_lower, dtype=float), np.asarray(self.xyz_upper, dtype=float)
            assert lo.shape == (3,) and hi.shape == (3,)
            if not np.all(lo < hi):
                raise ValueError("end_effector_box_constraint: xyz_lower must be below xyz_upper")
            d.ee_box_lower[:] = lo
            d.ee_box_upper[:] = hi

        Qd, Rd, Wd = np.diag(self.state_weight), np.diag(self.input_weight), np.diag(self.end_effector_weight)
        for M in (self.state_weight, self.input_weight, self.end_effector_weight):
            if np.abs(M - np.diag(np.diag(M))).max() > 0:
                raise NotImplementedError("only diagonal weights are supported")
        d.state_weight[: 3 * nq] = Qd
        d.input_weight[:nq] = Rd
        d.ee_weight[:] = Wd
        d.force_weight = float(self.balancing_settings.force_weight)
        d.xd[: 3 * nq] = self.xd[: 3 * nq]
        d.state_lb[: 3 * nq] = self.state_limit_lower
        d.state_ub[: 3 * nq] = self.state_limit_upper
        d.input_lb[:nq] = self.input_limit_lower
        d.input_ub[:nq] = self.input_limit_upper
        # controller_interface.cpp:330-356
        d.force_ub = 1e2
        d.force_lb = 0.0 if d.nf == 1 else -1e2

        index = {n: i for i, n in enumerate(names)}
        for i, n in enumerate(names):
            d.body_params[i][:] = self.balancing_settings.bodies[n].get_parameters()
        if d.balancing_enabled:
            for i, c in enumerate(self.balancing_settings.contacts):
                dc = d.contacts[i]
                dc.body1 = index.get(c.object1_name, -1)
                dc.body2 = index[c.object2_name]
                dc.mu = c.mu
                dc.r_co_o1[:] = c.r_co_o1
                dc.r_co_o2[:] = c.r_co_o2
                dc.normal[:] = c.normal
                dc.span[:] = np.asarray(c.span).ravel()

        if self.projectile_path_constraint_enabled and not d.obstacles_enabled:
            raise ValueError("projectile_path_constraint needs obstacles.enabled (the projectile is a dynamic obstacle)")
        if d.obstacles_enabled:
            spheres = list(self.chain.spheres)
            sidx = {s.name: i for i, s in enumerate(spheres)}
            for name, (pos, rad) in self.obstacle_settings.static_spheres.items():
                sidx[name] = len(spheres)
                spheres.append(robot.Sphere(name, -1, pos, rad))
            # the `ground` half-space z <= 0 is part of every collision model (add_ground_plane,
            # controller_interface.cpp:93-101,189); obstacles/dynamic.yaml:28-29 and sudden.yaml:29-30 pair against it
            sidx["ground"] = len(spheres)
            spheres.append(robot.Sphere("ground", -1, np.array([0.0, 0.0, 1.0]), 0.0, shape=B.UB_SHAPE_HALFSPACE))
            dyn = self.obstacle_settings.dynamic_obstacles
            if len(dyn) > B.UB_MAX_DYNAMIC_OBSTACLES:
                raise ValueError("too many dynamic obstacles")
            for j, o in enumerate(dyn):   # sphere riding on obstacle j: link = -2 - j, centre = its position state
                sidx[o.name] = len(spheres)
                spheres.append(robot.Sphere(o.name, -2 - j, o.modes[0].position, o.radius))
            d.n_dynamic_obstacles = len(dyn)
            used, pairs = {}, []
            def lookup(n):
                key = n[:-2] if n.endswith("_0") else n
                if key not in sidx:
                    raise ValueError(f"collision pair names an unknown collision object {n!r} (known: {sorted(sidx)})")
                return sidx[key]

            for a, b in self.obstacle_settings.collision_link_pairs:
                ia, ib = lookup(a), lookup(b)
                if spheres[ia].shape and not spheres[ib].shape:
                    ia, ib = ib, ia   # the sphere first, the half-space second
                for k in (ia, ib):
                    used.setdefault(k, len(used))
                pairs.append((used[ia], used[ib]))
            # ProjectilePathConstraint (controller_interface.cpp:272-294): one row per listed collision link, measured
            # from the origin of the link's frame = the centre of its collision sphere, to the LAST dynamic obstacle
            if self.projectile_path_constraint_enabled:
                links = self.projectile_path_collision_links
                if len(links) != len(self.projectile_path_distances):
                    raise RuntimeError("Number of distances and EEs must be equal!")   # projectile_path_constraint.h:57-62
                if not dyn:
                    raise ValueError("projectile_path_constraint needs a dynamic obstacle (the projectile)")
                if len(links) > B.UB_MAX_PROJECTILE_LINKS:
                    raise ValueError("too many projectile collision links")
                d.projectile_enabled, d.n_projectile_links = 1, len(links)
                d.projectile_scale = self.projectile_path_scale
                d.projectile_active = float(getattr(self, "projectile_active", 0.0))
                for i, n in enumerate(links):
                    k = lookup(n)
                    d.projectile_spheres[i] = used.setdefault(k, len(used))
                    d.projectile_distances[i] = float(self.projectile_path_distances[i])
            if len(used) > B.UB_MAX_SPHERES or len(pairs) > B.UB_MAX_PAIRS:
                raise ValueError("too many collision spheres / pairs")
            for k, slot in used.items():
                s = spheres[k]
                d.spheres[slot].link = s.link
                d.spheres[slot].shape = s.shape
                d.spheres[slot].radius = s.radius
                d.spheres[slot].offset[:] = s.offset
            for i, (a, b) in enumerate(pairs):
                d.pairs[i].a, d.pairs[i].b = a, b
            d.n_spheres, d.n_pairs = len(used), len(pairs)
            d.minimum_distance = float(self.obstacle_settings.minimum_distance)

        s = self.sqp.hpipm.slacks
        d.slacks.enabled = int(bool(s.enabled))
        d.slacks.input_box = int(bool(s.input_box))
        d.slacks.state_box = int(bool(s.state_box))
        d.slacks.poly_ineq = int(bool(s.poly_ineq))
        d.slacks.upper_L2_penalty = float(s.upper_L2_penalty)
        d.slacks.lower_L2_penalty = float(s.lower_L2_penalty)
        if s.enabled and (s.upper_L1_penalty or s.lower_L1_penalty or s.upper_low_bound or s.lower_low_bound):
            raise NotImplementedError("only pure L2 slack penalties with zero lower bound are supported")
        if s.enabled and s.upper_L2_penalty != s.lower_L2_penalty:
            raise NotImplementedError("upper/lower L2 slack penalties must be equal")

        # numerics: documented choices
        d.qp_method = 0
        d.rho_hard = 1.0e3
        d.qp_mu0, d.qp_thr0, d.qp_mu_target = 1.0e-1, 3.0, 1.0e-7
        d.qp_tol, d.reg_input = 1.0e-5, 1.0e-6
        d.alpha_decay, d.alpha_min = 0.5, 1e-4
        d.g_max, d.g_min, d.gamma_c, d.armijo_factor = 1e6, 1e-6, 1e-6, 1e-4
        d.delta_tol, d.cost_tol = float(self.sqp.delta_tol), float(self.sqp.cost_tol)
        return d
