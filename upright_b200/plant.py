"""Plants for the host-side closed loop: what the reference's simulation does to the things the controller only
observes.  Here: the dynamic obstacles of upright_sim (`upright_sim/src/upright_sim/simulation.py:300-435`,
configured under `simulation.dynamic_obstacles.obstacles`, e.g. `upright_cmd/config/obstacles/dynamic.yaml:38-75`)
and the in-flight gate of the projectile constraint (`upright_ros_interface/src/mrt_node.cpp:29-30,241-263`)."""
from __future__ import annotations

import numpy as np


class BallisticObstacles:
    """Uncontrolled dynamic obstacles for B instances: free flight under the current mode's acceleration, state
    reset to the next mode's initial values once its time has come (simulation.py:410-426: the check runs on the
    time at the START of a simulation step).  `relative` obstacles are placed relative to `offsets[b]`
    (the end-effector position at the start, simulation.py:337-343)."""

    def __init__(self, configs, batch, offsets=None):
        self.B = int(batch)
        self.obs = []
        for c in configs:
            if c.get("controlled", False):
                raise NotImplementedError("controlled (trajectory-tracking) obstacles are a simulator feature")
            off = np.zeros((self.B, 3))
            if c.get("relative", False) and offsets is not None:
                off = np.broadcast_to(np.asarray(offsets, dtype=float), (self.B, 3)).copy()
            modes = [dict(time=float(m["time"]), p=np.asarray(m["position"], dtype=float),
                          v=np.asarray(m["velocity"], dtype=float), a=np.asarray(m["acceleration"], dtype=float))
                     for m in c["modes"]]
            self.obs.append(dict(modes=modes, off=off, idx=0, r=None, v=None))
        self.start_time = None
        self.start(0.0)

    def __len__(self):
        return len(self.obs)

    def _enter(self, o, idx):
        m = o["modes"][idx]
        o["idx"] = idx
        o["r"] = m["p"][None, :] + o["off"]
        o["v"] = np.tile(m["v"], (self.B, 1))

    def start(self, t0=0.0):
        self.start_time = float(t0)
        for o in self.obs:
            self._enter(o, 0)

    def state(self):
        """[B, 9 n]: [r, v, a] per obstacle (simulation.py:623-633)."""
        if not self.obs:
            return np.zeros((self.B, 0))
        return np.hstack([np.hstack((o["r"], o["v"], np.tile(o["modes"][o["idx"]]["a"], (self.B, 1)))) for o in self.obs])

    def step(self, t, dt):
        """One simulation step starting at time t.  Returns True when an obstacle entered a new mode."""
        reset = False
        for o in self.obs:
            if o["idx"] < len(o["modes"]) - 1 and t - self.start_time >= o["modes"][o["idx"] + 1]["time"]:
                self._enter(o, o["idx"] + 1)
                reset = True
            a = o["modes"][o["idx"]]["a"][None, :]
            o["r"] = o["r"] + dt * o["v"] + 0.5 * dt * dt * a
            o["v"] = o["v"] + dt * a
        return reset


class ProjectileGate:
    """Pre-flight -> flight -> post-flight switch of mrt_node.cpp:241-263: the flag s of the target goes up once the
    projectile is above the activation height and down (for good) once it has dropped below the deactivation
    height.  Until the flight starts the controller keeps seeing the obstacle's nominal initial state."""

    PREFLIGHT, FLIGHT, POSTFLIGHT = 0, 1, 2

    def __init__(self, activation_height=1.0, deactivation_height=0.2):
        self.hi, self.lo = float(activation_height), float(deactivation_height)
        self.state = self.PREFLIGHT

    def update(self, z):
        if self.state == self.PREFLIGHT and z > self.hi:
            self.state = self.FLIGHT
        elif self.state == self.FLIGHT and z < self.lo:
            self.state = self.POSTFLIGHT
        return 1.0 if self.state == self.FLIGHT else 0.0

    @property
    def observing(self):
        return self.state != self.PREFLIGHT
