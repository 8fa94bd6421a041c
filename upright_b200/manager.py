"""Reference-facing control surface: `ControllerInterface`, `ControllerModel`,
`ControllerManager` with the method names and semantics of the reference
(`upright_control/src/pybindings.cpp:364-427`,
`upright_control/src/upright_control/manager.py:14-209`), backed by the CUDA
library through `engine.BatchedMPC`, plus `BatchedControllerManager` which
drives B instances at once (the receding-horizon shift, policy evaluation and
replan gate stay on the host in numpy, as they stay in Python in the
reference).
"""
from __future__ import annotations

import time

import numpy as np

from . import geometry as geo
from .engine import BatchedMPC
from .settings import ControllerSettings, TargetTrajectories
from .trajectory import StateInputTrajectory, interp_rows


class _RobotKinematics:
    """Host-side pose queries (role of `UprightRobotKinematics`, robot.py)."""

    def __init__(self, chain):
        self.chain = chain
        self._q = None

    def forward_xu(self, x, u=None):
        self._q = np.asarray(x)[: self.chain.nq]

    def link_pose(self):
        r, C = self.chain.tool_pose(self._q)
        return r, geo.rot_to_quat(C)


class ControllerModel:
    def __init__(self, settings):
        self.settings = settings
        self.robot = _RobotKinematics(settings.chain)
        self.geom = None

    @classmethod
    def from_config(cls, config, x0=None):
        return cls(ControllerSettings(config=config, x0=x0))

    def update(self, x, u=None):
        self.robot.forward_xu(x, u)


class ControllerInterface:
    """Single-robot MPC object (batch of one) with the pybind method surface."""

    def __init__(self, settings, precision="f32"):
        self.settings = settings
        self.desc = settings.to_desc()
        self._engine = BatchedMPC(self.desc, precision)
        self._core = _RecedingHorizon(self._engine, settings, batch=1)
        self._last_ms = 0.0

    def getStateDim(self):
        return self._engine.nx

    def getInputDim(self):
        return self._engine.nu

    def setObservation(self, t, x, u):
        self._core.observe(t, np.asarray(x, dtype=float)[None, :])

    def setTargetTrajectories(self, targets):
        self._core.targets = [targets]

    def reset(self, targets):
        self._core.reset([targets])

    def advanceMpc(self):
        self._core.advance()

    def getLastSolveTime(self):
        return self._engine.last_solve_ms()

    def getMpcSolution(self, ts, xs, us):
        """Out-parameters like the bound std::vectors: python lists are extended."""
        T, X, U = self._core.solution()
        ts.extend(T.tolist())
        xs.extend(list(X[0]))
        us.extend(list(np.vstack((U[0], U[0][-1:]))))

    def evaluateMpcSolution(self, current_time, current_state, opt_state, opt_input):
        x, u = self._core.evaluate(current_time, np.asarray(current_state, dtype=float)[None, :])
        opt_state[:] = x[0]
        opt_input[:] = u[0]

    def getLinearFeedbackGain(self, t):
        return self._core.gain(t)[0]

    def getBias(self, t):
        return self._core.bias(t)[0]

    def cost(self, t, x, u):
        tgt = self._core.targets[0].get_desired_state(t)[:3]
        return float(self._engine.eval("cost", x, u, target=tgt[None])[0, 0])

    def getCostValue(self, name, t, x, u):
        if name == "inertial_alignment_cost":   # controller_interface.cpp:296-305
            return float(self._engine.eval(name, x, u)[0, 0])
        if name not in ("state_input_cost", "end_effector_cost"):
            raise RuntimeError(f"unknown cost {name}")
        total = self.cost(t, x, u)
        zero_w = self._engine.eval("cost", x, u)[0, 0]  # without the EE term
        if name == "end_effector_cost":
            return float(total - zero_w)
        # the quadratic state-input term alone (controller_interface.cpp:136): the "cost" probe also carries the
        # inertial-alignment cost when that is enabled
        ia = self._engine.eval("inertial_alignment_cost", x, u)[0, 0] if self.desc.ia_cost_enabled else 0.0
        return float(zero_w - ia)

    # ---- model queries of ocs2::PythonInterface bound at pybindings.cpp:388-397
    def flowMap(self, t, x, u):
        """x' = [v, a, jerk] for the robot, [v_o, a_o, 0] per dynamic obstacle (dynamics/system_dynamics.h:15-38,86-104);
        the contact forces do not enter."""
        x, u = np.asarray(x, dtype=float), np.asarray(u, dtype=float)
        nq, no = self.desc.nq, self.desc.n_dynamic_obstacles if self.desc.obstacles_enabled else 0
        out = np.concatenate((x[nq:3 * nq], u[:nq]))
        for j in range(no):
            o = x[3 * nq + 9 * j: 3 * nq + 9 * (j + 1)]
            out = np.concatenate((out, o[3:9], np.zeros(3)))
        return out

    def flowMapLinearApproximation(self, t, x, u):
        """-> object with f, dfdx, dfdu (ocs2::VectorFunctionLinearApproximation): the flow map is linear."""
        nq, nx, nu = self.desc.nq, self.getStateDim(), self.getInputDim()
        A, Bm = np.zeros((nx, nx)), np.zeros((nx, nu))
        A[:2 * nq, nq:3 * nq] = np.eye(2 * nq)
        Bm[2 * nq:3 * nq, :nq] = np.eye(nq)
        for o in range(3 * nq, nx, 9):
            A[o:o + 6, o + 3:o + 9] = np.eye(6)
        return _Approximation(f=self.flowMap(t, x, u), dfdx=A, dfdu=Bm)

    def costQuadraticApproximation(self, t, x, u):
        """-> object with f, dfdx, dfdu, dfdxx, dfdux, dfduu of the intermediate cost: quadratic state-input cost
        (quadratic_joint_state_input_cost.h:9-33) + end-effector cost with its Gauss-Newton Hessian
        (end_effector_cost.h:48-84) [+ the inertial-alignment Gauss-Newton cost is not part of this query]."""
        d = self.desc
        x, u = np.asarray(x, dtype=float), np.asarray(u, dtype=float)
        nq, nxr, nx, nu = d.nq, 3 * d.nq, self.getStateDim(), self.getInputDim()
        Q = np.zeros(nx)
        Q[:nxr] = np.array(d.state_weight[:nxr])
        xd = np.zeros(nx)
        xd[:nxr] = np.array(d.xd[:nxr])
        Rw = np.concatenate((np.array(d.input_weight[:nq]), np.full(nu - nq, d.force_weight)))
        W = np.array(d.ee_weight[:3])
        tgt = self._core.targets[0].get_desired_state(t)[:3]
        r = self._engine.eval("end_effector_position", x, u)[0]
        J = np.zeros((3, nx))
        J[:, :nq] = self._engine.eval("end_effector_jacobian", x, u)[0].reshape(3, nq)
        e = r - tgt
        f = 0.5 * Q @ (x - xd) ** 2 + 0.5 * Rw @ u ** 2 + 0.5 * W @ e ** 2
        return _Approximation(f=float(f), dfdx=Q * (x - xd) + J.T @ (W * e), dfdu=Rw * u,
                              dfdxx=np.diag(Q) + J.T @ (W[:, None] * J), dfdux=np.zeros((nu, nx)), dfduu=np.diag(Rw))

    def getStateInputEqualityConstraintValue(self, name, t, x, u):
        if name != "object_dynamics":
            raise RuntimeError(f"unknown equality constraint {name}")
        return self._engine.eval("object_dynamics", x, u)[0]

    def getStateInputInequalityConstraintValue(self, name, t, x, u):
        # names as registered at controller_interface.cpp:221,266,291,312,347
        if name not in ("contact_forces", "obstacle_avoidance", "end_effector_box_constraint",
                        "inertial_alignment_constraint", "projectile_constraint"):
            raise RuntimeError(f"unknown inequality constraint {name}")
        tgt = None
        if name == "end_effector_box_constraint":
            tgt = self._core.targets[0].get_desired_state(t)[:3][None]
        return self._engine.eval(name, x, u, target=tgt)[0]


class _Approximation:
    """Attribute bag with the field names of ocs2's Vector / ScalarFunction*Approximation bindings."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class BalancingConstraintWrapper:
    """`bindings.BalancingConstraintWrapper` (balancing_constraint_wrapper.h:15-66; used by
    upright_cmd/scripts/misc/balance_at_given_configuration.py:26): getLinearApproximation(t, x, u) -> f, dfdx with the
    contact-force rows first and the object-dynamics rows behind them (`approx.f << a.f, b.f`); as in the reference
    only the state Jacobian is carried (`approx(nv, state.size(), 0)`)."""

    def __init__(self, settings, precision="f64"):
        if not settings.balancing_settings.enabled:
            raise RuntimeError("Balancing settings not enabled.")
        self.desc = settings.to_desc()
        self._engine = BatchedMPC(self.desc, precision)

    def getLinearApproximation(self, t, x, u):
        e = self._engine
        nxr, nu = e.nx_robot, e.nu
        x, u = np.asarray(x, dtype=float), np.asarray(u, dtype=float)
        a_f = e.eval("contact_forces", x, u)[0] if self.desc.n_fric else np.zeros(0)
        b_f = e.eval("object_dynamics", x, u)[0]
        jac = e.eval("object_dynamics_jacobian", x, u)[0].reshape(e.n_eq, nxr + nu)
        dfdx = np.zeros((a_f.size + b_f.size, e.nx))
        dfdx[a_f.size:, :nxr] = jac[:, :nxr]
        return _Approximation(f=np.concatenate((a_f, b_f)), dfdx=dfdx, dfdu=np.zeros((a_f.size + b_f.size, 0)),
                              dfdu_dynamics=jac[:, nxr:])


class _RecedingHorizon:
    """Receding-horizon bookkeeping for B instances: observation, warm-start
    shift of the previous solution onto the new time grid, policy evaluation.
    Semantics of ocs2 MPC_MRT_Interface as used at manager.py:156-176 [EXT]."""

    def __init__(self, engine, settings, batch):
        self.engine, self.settings, self.B = engine, settings, batch
        self.dt = settings.sqp.dt
        self.N = engine.N
        self.targets = None
        self.t_obs, self.x_obs = 0.0, None
        self.t0 = None
        self.X = self.U = self.K = None
        self.first = True
        self.status = None
        self.stats = None
        self.use_feedback = bool(settings.sqp.use_feedback_policy)
        self.body_params = None

    def reset(self, targets):
        self.targets = targets
        self.X = self.U = self.K = None
        self.t0 = None
        self.first = True

    def observe(self, t, x):
        self.t_obs, self.x_obs = float(t), np.array(x, dtype=float)

    def _knot_targets(self, t):
        """Desired end-effector position at every knot — and, when the orientation part of the end-effector weight is
        non-zero, the desired quaternion behind it (interpolate_end_effector_pose, reference_trajectory.h:18-47)."""
        times = t + self.dt * np.arange(self.N + 1)
        ts = self.engine.target_stride if hasattr(self.engine, "target_stride") else 3
        tg = np.empty((self.B, self.N + 1, ts))
        for b in range(self.B):
            tt = self.targets[b if len(self.targets) > 1 else 0]
            tg[b] = tt.positions_at(times) if ts == 3 else tt.poses_at(times)
        return tg

    def operating_guess(self, t):
        """Initial guess from the operating trajectory (ocs2::OperatingPoints [EXT], chosen at
        controller_interface.cpp:380-387): u_k = inputs(t_k), x_{k+1} = states(t_{k+1}), linear interpolation clamped
        at both ends; x_0 is the observation."""
        from .trajectory import interp_rows
        s = self.settings
        times = t + self.dt * np.arange(self.N + 1)
        ts, xs, us = np.asarray(s.operating_times), np.asarray(s.operating_states), np.asarray(s.operating_inputs)
        X1 = np.stack([interp_rows(ts, xs, tk) for tk in times])
        U1 = np.stack([interp_rows(ts, us, tk) for tk in times[:-1]])
        X = np.tile(X1, (self.B, 1, 1))
        X[:, 0] = self.x_obs
        return X, np.tile(U1, (self.B, 1, 1))

    def advance(self):
        t = self.t_obs
        warm = (self.X is not None) and not self.settings.mpc.cold_start
        X = U = None
        operating = bool(getattr(self.settings, "use_operating_points", False))
        if operating and not warm:
            # the initializer covers the whole horizon; handed to the solver as its starting iterate
            X, U = self.operating_guess(t)
        if warm:
            # previous primal solution interpolated at the new grid, tail held
            told = self.t0 + self.dt * np.arange(self.N + 1)
            tnew = t + self.dt * np.arange(self.N + 1)
            X = np.empty_like(self.X)
            U = np.empty_like(self.U)
            Uext = np.concatenate((self.U, self.U[:, -1:, :]), axis=1)
            for k in range(self.N + 1):
                i = int(np.clip(np.searchsorted(told, tnew[k], side="right") - 1, 0, self.N - 1))
                w = np.clip((tnew[k] - told[i]) / self.dt, 0.0, 1.0)
                X[:, k] = (1 - w) * self.X[:, i] + w * self.X[:, i + 1]
                if k < self.N:
                    U[:, k] = (1 - w) * Uext[:, i] + w * Uext[:, i + 1]
            if operating:   # knots beyond the previous horizon come from the initializer, not from holding the tail
                Xo, Uo = self.operating_guess(t)
                X[:, tnew > told[-1] + 1e-12] = Xo[:, tnew > told[-1] + 1e-12]
                U[:, tnew[:-1] > told[-1] - 1e-12] = Uo[:, tnew[:-1] > told[-1] - 1e-12]   # intervals not covered
        iters = self.settings.sqp.init_sqp_iteration if self.first else self.settings.sqp.sqp_iteration
        self.engine.set_option("sqp_iteration", int(iters))
        if getattr(self.settings, "projectile_path_constraint_enabled", False):
            # the flag s = last element of the FIRST target state (projectile_path_constraint.h:78-80), raised by the
            # caller while the projectile is in flight (mrt_node.cpp:241-263); one flag for the whole batch
            self.engine.set_option("projectile_active", int(self.targets[0].xs[0][7] > 0.5))
        out = self.engine.solve(self.x_obs, self._knot_targets(t), self.body_params, X=X, U=U, warm=X is not None,
                                want_gains=self.use_feedback, rescue=True)
        self.X, self.U = out["X"], out["U"]
        self.K = out.get("K")
        self.status, self.stats = out["status"], out["stats"]
        self.t0 = t
        self.first = False
        if np.any(self.status == 3):
            raise RuntimeError("MPC solve produced non-finite values")

    def solution(self):
        return self.t0 + self.dt * np.arange(self.N + 1), self.X, self.U

    def _interp(self, t):
        s = (t - self.t0) / self.dt
        i = int(np.clip(np.floor(s), 0, self.N - 1))
        w = float(np.clip(s - i, 0.0, 1.0))
        return i, w

    def gain(self, t):
        if self.K is None:
            return np.zeros((self.B, self.engine.nu, getattr(self.engine, "nx_robot", self.engine.nx)))
        i, w = self._interp(t)
        j = min(i + 1, self.N - 1)
        return (1 - w) * self.K[:, i] + w * self.K[:, j]

    def evaluate(self, t, x):
        i, w = self._interp(t)
        x_nom = (1 - w) * self.X[:, i] + w * self.X[:, i + 1]
        j = min(i + 1, self.N - 1)
        u_ff = (1 - w) * self.U[:, i] + w * self.U[:, j]
        if self.use_feedback and self.K is not None:
            nxr = self.K.shape[-1]   # gains act on the robot state (dynamic-obstacle states are uncontrolled)
            u = u_ff + np.einsum("bij,bj->bi", self.gain(t), (x - x_nom)[:, :nxr])
        else:
            u = u_ff
        return x_nom, u

    def bias(self, t):
        x_nom, _ = self.evaluate(t, np.zeros((self.B, self.engine.nx)))
        i, w = self._interp(t)
        j = min(i + 1, self.N - 1)
        u_ff = (1 - w) * self.U[:, i] + w * self.U[:, j]
        return u_ff - np.einsum("bij,bj->bi", self.gain(t), x_nom[:, : self.gain(t).shape[-1]])


class ControllerManager:
    """manager.py:100-209 — replan gate + policy evaluation for one robot."""

    def __init__(self, model, ref_trajectory, timestep, precision="f32"):
        self.model = model
        self.ref = ref_trajectory
        self.timestep = timestep
        self.mpc = ControllerInterface(self.model.settings, precision)
        self.mpc.reset(self.ref)
        self.last_planning_time = -np.inf
        self.x_opt = np.zeros(self.model.settings.dims.x())
        self.u_opt = np.zeros(self.model.settings.dims.u())
        self.replanning_times = []
        self.replanning_durations = []

    @classmethod
    def from_config(cls, config, x0=None, precision="f32"):
        model = ControllerModel.from_config(config, x0=x0)
        timestep = config["tracking"]["min_policy_update_time"]
        model.update(x=model.settings.initial_state)
        r_ew_w, Q_we = model.robot.link_pose()
        ref = TargetTrajectories.from_config(config, r_ew_w, Q_we, np.zeros(model.settings.dims.u()))
        return cls(model, ref, timestep, precision)

    def update(self, ref):
        self.ref = ref
        self.mpc.reset(self.ref)

    def warmstart(self):
        x0 = self.model.settings.initial_state
        self.mpc.setObservation(0, x0, np.zeros(self.model.settings.dims.u()))
        self.mpc.advanceMpc()
        self.last_planning_time = 0

    def step(self, t, x):
        self.mpc.setObservation(t, x, self.u_opt)
        if t >= self.last_planning_time + self.timestep:
            t0 = time.time()
            self.mpc.advanceMpc()
            t1 = time.time()
            self.last_planning_time = t
            self.replanning_times.append(t)
            self.replanning_durations.append(t1 - t0)
        self.mpc.evaluateMpcSolution(t, x, self.x_opt, self.u_opt)
        return self.x_opt, self.u_opt

    def get_mpc_trajectory(self):
        ts, xs, us = [], [], []
        self.mpc.getMpcSolution(ts, xs, us)
        return np.array(ts), np.array(xs), np.array(us)

    def plan(self, timestep, duration):
        ts, xs, us = [], [], []
        t = 0.0
        x = self.model.settings.initial_state
        while t <= duration:
            x, u = self.step(t, x)
            ts.append(t)
            xs.append(x.copy())
            us.append(u.copy())
            t += timestep
        return StateInputTrajectory(ts, xs, us)


class BatchedControllerManager:
    """B robots / scenarios at once: same `step(t, x)` contract with x [B, nx]."""

    def __init__(self, settings, targets, timestep=None, body_params=None, precision="f32"):
        self.settings = settings
        self.desc = settings.to_desc()
        self.engine = BatchedMPC(self.desc, precision)
        self.B = len(targets)
        self.core = _RecedingHorizon(self.engine, settings, self.B)
        self.core.reset(list(targets))
        self.core.body_params = body_params
        self.timestep = settings.tracking.min_policy_update_time if timestep is None else timestep
        self.last_planning_time = -np.inf
        self.replanning_times, self.replanning_durations = [], []

    def step(self, t, x):
        x = np.asarray(x, dtype=float)
        self.core.observe(t, x)
        if t >= self.last_planning_time + self.timestep:
            t0 = time.time()
            self.core.advance()
            self.replanning_durations.append(time.time() - t0)
            self.replanning_times.append(t)
            self.last_planning_time = t
        return self.core.evaluate(t, x)

    def get_mpc_trajectory(self):
        return self.core.solution()

    def rollout(self, x0, duration, sim_timestep, log_stride=1, log=True, obstacles=None, obstacle_offsets=None):
        """Closed loop for `duration` seconds entirely on the device (`ub_closed_loop`): the loop of
        `mpc_sim.py:118-160` — `step(t, x)`, `u_cmd = Kx (xd - x) + u`, integrate — for all B robots, with the
        model's triple integrator as the plant.  The waypoint times must be shared by all targets.

        With dynamic obstacles x0 carries their start states behind the robot state and `obstacles` their simulated
        mode schedules (`simulation.dynamic_obstacles.obstacles[*].modes`, free flight + resets: the device-side
        counterpart of `plant.BallisticObstacles`); the projectile gate of the ROS node stays a host matter
        (`rollout_host`)."""
        tr = self.settings.tracking
        tt = self.core.targets[0].ts
        for tg in self.core.targets:
            if list(tg.ts) != list(tt):
                raise ValueError("rollout needs the same waypoint times for every instance")
        pos = np.array([[x[:3] for x in tg.xs] for tg in self.core.targets])
        if pos.shape[0] == 1 and self.B > 1:
            pos = np.repeat(pos, self.B, axis=0)
        n_steps = int(round(duration / sim_timestep))
        return self.engine.closed_loop(
            x0, tt, pos, n_steps, sim_timestep, self.timestep, body_params=self.core.body_params,
            use_feedback=bool(self.settings.sqp.use_feedback_policy), cold_start=bool(self.settings.mpc.cold_start),
            init_sqp_iteration=self.settings.sqp.init_sqp_iteration, sqp_iteration=self.settings.sqp.sqp_iteration,
            gains=(getattr(tr, "kp", 0.0), getattr(tr, "kv", 0.0), getattr(tr, "ka", 0.0)), log_stride=log_stride, log=log,
            obstacles=obstacles, obstacle_offsets=obstacle_offsets)

    def rollout_host(self, x0, duration, sim_timestep, obstacles=None, gate=None, log_stride=1):
        """The loop of `mpc_sim.py:118-160` on the host, one `step(t, x)` per simulation step for all B robots, for
        the problems the device loop does not take: dynamic obstacles.  `x0` is the robot state [B, 3 nq];
        `obstacles` a `plant.BallisticObstacles` (its state fills the obstacle columns of x every step, as
        `env.dynamic_obstacle_state()` does at mpc_sim.py:120-121), `gate` a `plant.ProjectileGate` driven by the
        height of the LAST obstacle of instance 0 (one flag s for the batch: it is a problem option).  The robot
        plant is the model's exact triple integrator, as in `ub_closed_loop`."""
        nq = self.settings.dims.robot.q
        nxr = 3 * nq
        tr = self.settings.tracking
        kp, kv, ka = (getattr(tr, k, 0.0) for k in ("kp", "kv", "ka"))
        xr = np.array(x0, dtype=float).reshape(self.B, nxr)
        n_obs = self.desc.n_dynamic_obstacles
        if n_obs and (obstacles is None or len(obstacles) != n_obs):
            raise ValueError(f"the problem carries {n_obs} dynamic obstacle(s): pass a matching obstacle plant")
        nominal = obstacles.state() if n_obs else np.zeros((self.B, 0))
        n_steps = int(round(duration / sim_timestep))
        xs, us, flags = [], [], []
        h = sim_timestep
        for step in range(n_steps):
            t = h * step
            xo = obstacles.state() if n_obs else nominal
            if gate is not None:
                s = gate.update(float(xo[0, -7]))          # z of the last obstacle
                for tg in self.core.targets:
                    tg.xs[0][7] = s
                if not gate.observing:
                    xo = nominal                           # mrt_node.cpp:265-270: state updated once past pre-flight
                flags.append(s)
            x = np.hstack((xr, xo))
            xd, u = self.step(t, x)
            e = (xd - x)[:, :nxr]
            ucmd = kp * e[:, :nq] + kv * e[:, nq:2 * nq] + ka * e[:, 2 * nq:] + u[:, :nq]
            if step % log_stride == 0:
                xs.append(x.copy())
                us.append(ucmd.copy())
            q, v, a = xr[:, :nq], xr[:, nq:2 * nq], xr[:, 2 * nq:]
            xr = np.hstack((q + h * v + 0.5 * h * h * a + h**3 / 6 * ucmd, v + h * a + 0.5 * h * h * ucmd, a + h * ucmd))
            if n_obs:
                obstacles.step(t, h)
        xo = obstacles.state() if n_obs else nominal
        return dict(xs=np.stack(xs, 1), us=np.stack(us, 1), x_final=np.hstack((xr, xo)),
                    n_replans=len(self.replanning_times), flags=np.array(flags))
