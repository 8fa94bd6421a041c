"""Small trajectory helpers with the names `upright_control.trajectory` exposes
(upright_control/src/upright_control/trajectory.py:7-75): the double-integrator
plant `mpc_sim.py:148-155` steps the robot with, the generic state-input
trajectory container and the (x,u) <-> (q,v,a) mapping."""
from __future__ import annotations

import numpy as np


class DoubleIntegrator:
    """v' = a, a' = u, integrated exactly over dt (trajectory.py:7-32)."""

    def __init__(self, n):
        self.n = n

    def integrate(self, v, a, u, dt):
        return v + dt * a + 0.5 * dt * dt * u, a + dt * u

    def integrate_approx(self, v, a, u, dt):
        a_new = a + dt * u
        return v + dt * a_new, a_new


class StateInputTrajectory:
    def __init__(self, ts, xs, us):
        assert len(ts) == len(xs) == len(us)
        self.ts, self.xs, self.us = ts, xs, us

    @classmethod
    def load(cls, filename):
        with np.load(filename) as data:
            return cls(ts=data["ts"], xs=data["xs"], us=data["us"])

    def save(self, filename):
        np.savez_compressed(filename, ts=self.ts, xs=self.xs, us=self.us)

    def __getitem__(self, idx):
        return self.ts[idx], self.xs[idx], self.us[idx]

    def __len__(self):
        return len(self.ts)


class StateInputMapping:
    def __init__(self, dims):
        self.dims = dims

    def xu2qva(self, x, u=None):
        q = x[: self.dims.q]
        v = x[self.dims.q: self.dims.q + self.dims.v]
        a = x[self.dims.q + self.dims.v: self.dims.q + 2 * self.dims.v]
        return q, v, a

    def qva2xu(self, q, v, a):
        return np.concatenate((q, v, a)), None


def interp_rows(ts, rows, t):
    """Linear interpolation of row-stacked samples at scalar time t, clamped at
    both ends (ocs2::LinearInterpolation semantics)."""
    ts = np.asarray(ts)
    if t <= ts[0]:
        return rows[0].copy()
    if t >= ts[-1]:
        return rows[-1].copy()
    i = int(np.searchsorted(ts, t, side="right")) - 1
    w = (t - ts[i]) / (ts[i + 1] - ts[i])
    return (1 - w) * rows[i] + w * rows[i + 1]
