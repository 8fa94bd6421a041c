"""Batched MPC engine: thin Python layer over the C-ABI CUDA library.

`BatchedMPC.solve()` is the host-buffer path (numpy float64 in/out, copies
inside the call — the drop-in replacement of B x `advanceMpc()`);
`BatchedMPC.solve_device()` keeps everything in torch CUDA tensors (f32) and
only enqueues the kernel on the current stream.  torch is plumbing here
(device memory, streams); all arithmetic is in `csrc/`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import bindings as B

LAYOUT_FIELDS = ("Z", "DZ", "GAP", "LG", "LC", "LR", "LJP", "LHO", "LJO", "DF", "RHOE", "YE", "RHOT", "YT", "TL", "DD",
                 "GP", "VE", "FAC", "WF", "FBB", "XN", "UN", "XW", "UW", "TG", "BD", "LIA", "LJA", "XO", "DXO", "TGQ", "LRO", "LJQ", "total",
                 "bG", "bL", "bD", "bGl", "bQ", "bsize", "sM", "sP", "sPv", "sFB", "sGf", "sFv", "sFl", "sVec", "sDst", "sDxn",
                 "sRv", "sScr", "sDFC", "sUS", "sCst", "sTL", "sDD", "sSmZ", "sSmX", "sSmU", "sSmJ", "sSmW", "sBar", "s_total", "rw")
# workspace blocks that hold doubles whatever the kernels' matrix type (ub_solver.cuh: compute_layout)
LAYOUT_DOUBLE_BLOCKS = ("Z", "GAP", "LG", "LR", "RHOE", "YE", "RHOT", "YT", "TL", "GP", "VE")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BatchedMPC:
    """One immutable problem configuration on one GPU."""

    def __init__(self, desc: B.ProblemDesc, precision: str = "f32"):
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        self.lib = B.load_library()
        self.lib.ub_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        self.lib.ub_workspace_layout.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32)]
        self.desc = desc
        self.precision = precision
        self.flags = B.UB_COMPUTE_F64 if precision == "f64" else 0
        handle = C.c_void_p()
        B.check(self.lib.ub_problem_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        dims = (C.c_int32 * 8)()
        B.check(self.lib.ub_problem_dims(self.handle, dims))
        self.nx, self.nu, self.n_eq, self.n_ineq, self.n_term, self.N, self.nb, self.nc = list(dims)
        # nx counts the dynamic-obstacle states too (9 each, behind the robot state); gains act on the robot state
        self.nx_robot = 3 * desc.nq
        # columns of a target row: desired position (3), + desired quaternion [x y z w] when the orientation part of
        # the end-effector weight is non-zero (7)
        self.target_stride = 7 if any(desc.ee_weight[i] != 0 for i in (3, 4, 5)) else 3
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ub_problem_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------- host path
    def solve(self, x0, target, body_params=None, X=None, U=None, warm=False, want_gains=False, out=None, rescue=False):
        """numpy float64 host buffers; returns dict(X, U, status, stats[, K]).

        `out`: a dict returned by an earlier call with the same batch size — its arrays are overwritten and
        returned again (what a C caller does with its own buffers; saves the page faults of fresh arrays).
        `rescue`: instances the fp32 kernels end with status NAN are solved again by the fp64 kernels
        (`UB_RESCUE_F64`)."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        Bn = x0.shape[0]
        assert x0.shape == (Bn, self.nx)
        target = np.ascontiguousarray(np.asarray(target, dtype=np.float64).reshape(Bn, self.N + 1, self.target_stride))
        if body_params is not None:
            body_params = np.ascontiguousarray(body_params, dtype=np.float64).reshape(Bn, self.nb, 10)
        reuse = out is not None and out["X"].shape == (Bn, self.N + 1, self.nx) and (("K" in out) == bool(want_gains))
        if warm:
            Xw = np.ascontiguousarray(X, dtype=np.float64).reshape(Bn, self.N + 1, self.nx)
            Uw = np.ascontiguousarray(U, dtype=np.float64).reshape(Bn, self.N, self.nu)
            if reuse:
                X, U = out["X"], out["U"]
                if X is not Xw:
                    X[...] = Xw
                if U is not Uw:
                    U[...] = Uw
            else:
                X, U = Xw.copy(), Uw.copy()
        elif reuse:
            X, U = out["X"], out["U"]
        else:
            X = np.empty((Bn, self.N + 1, self.nx))
            U = np.empty((Bn, self.N, self.nu))
        if reuse:
            K, status, stats = out.get("K"), out["status"], out["stats"]
        else:
            K = np.empty((Bn, self.N, self.nu, self.nx_robot)) if want_gains else None
            status = np.empty(Bn, dtype=np.int32)
            stats = np.empty((Bn, B.UB_STATS))
        flags = self.flags | (B.UB_WARM_START if warm else 0) | (B.UB_RESCUE_F64 if rescue else 0)
        B.check(self.lib.ub_solve_batch(self.handle, Bn, _ptr(x0), _ptr(target), _ptr(body_params), _ptr(X), _ptr(U),
                                        _ptr(K), _ptr(status), _ptr(stats), None, 0, flags, None))
        res = dict(X=X, U=U, status=status, stats=stats)
        if want_gains:
            res["K"] = K
        return res

    # ----------------------------------------------------------- device path
    @property
    def torch_dtype(self):
        import torch
        return torch.float64 if self.precision == "f64" else torch.float32

    def workspace(self, Bn, device=None):
        import torch
        nbytes = self.lib.ub_workspace_bytes(self.handle, Bn, self.flags)
        if self._ws is None or self._ws.numel() < nbytes or (device is not None and self._ws.device != device):
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device or "cuda")
        return self._ws

    def solve_device(self, x0, target, body_params=None, X=None, U=None, K=None, status=None, stats=None, warm=False):
        """torch CUDA tensors of `self.torch_dtype`; enqueues on the current stream, no sync."""
        import torch
        Bn = x0.shape[0]
        dt, dev = self.torch_dtype, x0.device
        assert x0.dtype == dt and target.dtype == dt and x0.is_contiguous() and target.is_contiguous()
        if X is None:
            assert not warm
            X = torch.empty((Bn, self.N + 1, self.nx), dtype=dt, device=dev)
            U = torch.empty((Bn, self.N, self.nu), dtype=dt, device=dev)
        if status is None:
            status = torch.empty(Bn, dtype=torch.int32, device=dev)
        if stats is None:
            stats = torch.empty((Bn, B.UB_STATS), dtype=dt, device=dev)
        ws = self.workspace(Bn, dev)
        flags = self.flags | B.UB_PTRS_DEVICE | (B.UB_WARM_START if warm else 0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        B.check(self.lib.ub_solve_batch(self.handle, Bn, p(x0), p(target), p(body_params), p(X), p(U), p(K),
                                        p(status), p(stats), p(ws), ws.numel(), flags, C.c_void_p(stream)))
        return dict(X=X, U=U, K=K, status=status, stats=stats)

    # ----------------------------------------------------------- closed loop
    def closed_loop(self, x0, target_times, target_pos, n_steps, sim_dt, replan_period, body_params=None,
                    use_feedback=True, cold_start=False, init_sqp_iteration=1, sqp_iteration=1, gains=(0.0, 0.0, 0.0),
                    log_stride=1, log=True, obstacles=None, obstacle_offsets=None):
        """B closed-loop rollouts on the device (`ub_closed_loop`): replan gate, warm-start shift, policy
        evaluation and the triple-integrator plant as in `mpc_sim.py:118-160`.

        x0 [B, nx]; target_times [M]; target_pos [B, M, 3].  Returns dict(xs [B, n_log, nx], us [B, n_log, nq],
        x_final [B, nx], n_replans, status_counts [B, 4]).

        Problems with dynamic obstacles (nx = 3 nq + 9 per obstacle): `obstacles` = one list of modes per obstacle,
        each mode a dict(time, position, velocity, acceleration) as under `simulation.dynamic_obstacles.obstacles`
        (obstacles/dynamic.yaml:38-75); `obstacle_offsets` [B, n, 3] places `relative` obstacles per instance.  The
        obstacle columns of x0 are their states at the start."""
        nobs = self.desc.n_dynamic_obstacles if self.desc.obstacles_enabled else 0
        if obstacles is not None:
            if len(obstacles) != nobs:
                raise ValueError(f"the problem carries {nobs} dynamic obstacle(s)")
            nm = (C.c_int32 * nobs)(*[len(m) for m in obstacles])
            modes = (B.ObstacleMode * (nobs * B.UB_MAX_OBSTACLE_MODES))()
            for j, ms in enumerate(obstacles):
                if not 1 <= len(ms) <= B.UB_MAX_OBSTACLE_MODES:
                    raise ValueError("1 .. UB_MAX_OBSTACLE_MODES modes per obstacle")
                for m, md in enumerate(ms):
                    e = modes[j * B.UB_MAX_OBSTACLE_MODES + m]
                    e.time = float(md["time"])
                    e.position[:] = [float(v) for v in md["position"]]
                    e.velocity[:] = [float(v) for v in md["velocity"]]
                    e.acceleration[:] = [float(v) for v in md["acceleration"]]
            off = None if obstacle_offsets is None else np.ascontiguousarray(obstacle_offsets, dtype=np.float64).reshape(-1, nobs, 3)
            B.check(self.lib.ub_closed_loop_set_obstacles(self.handle, nobs, nm, modes, 0 if off is None else off.shape[0], _ptr(off)))
        else:
            B.check(self.lib.ub_closed_loop_set_obstacles(self.handle, 0, None, None, 0, None))
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        Bn = x0.shape[0]
        tt = np.ascontiguousarray(np.atleast_1d(target_times), dtype=np.float64)
        M = tt.shape[0]
        tp = np.ascontiguousarray(np.asarray(target_pos, dtype=np.float64).reshape(Bn, M, 3))
        bp = None if body_params is None else np.ascontiguousarray(body_params, dtype=np.float64).reshape(Bn, self.nb, 10)
        prm = B.ClosedLoopParams(float(sim_dt), float(replan_period), int(n_steps), int(log_stride), int(bool(use_feedback)),
                                 int(bool(cold_start)), int(init_sqp_iteration), int(sqp_iteration), float(gains[0]),
                                 float(gains[1]), float(gains[2]))
        nq = self.desc.nq
        n_log = (int(n_steps) + int(log_stride) - 1) // int(log_stride) if log else 0
        xs = np.empty((Bn, n_log, self.nx)) if log else None
        us = np.empty((Bn, n_log, nq)) if log else None
        xf = np.empty((Bn, self.nx))
        counts = np.zeros((Bn, 4), dtype=np.int32)
        nrep = C.c_int32()
        B.check(self.lib.ub_closed_loop(self.handle, Bn, _ptr(x0), _ptr(tt), _ptr(tp), M, _ptr(bp), C.byref(prm), _ptr(xs),
                                        _ptr(us), _ptr(xf), C.byref(nrep), _ptr(counts), self.flags, None))
        return dict(xs=xs, us=us, x_final=xf, n_replans=int(nrep.value), status_counts=counts)

    def last_solve_ms(self):
        return float(self.lib.ub_last_solve_ms(self.handle))

    # -------------------------------------------------------- probes / debug
    def eval(self, name, x, u, target=None, body_params=None):
        """Named constraint/cost probes of the reference interface, batched."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        M = x.shape[0]
        cap = M * max(self.n_eq * (self.nx_robot + self.nu), 5 * self.nc, self.desc.n_pairs, self.desc.n_projectile_links,
                      3 * self.desc.nq, 6, 1)
        out = np.zeros(cap)
        rows = C.c_int32()
        tg = None if target is None else np.ascontiguousarray(target, dtype=np.float64).reshape(M, 3)
        bp = None if body_params is None else np.ascontiguousarray(body_params, dtype=np.float64)
        B.check(self.lib.ub_eval(self.handle, name.encode(), M, _ptr(x), _ptr(u), _ptr(tg), _ptr(bp), _ptr(out), cap,
                                 C.byref(rows)))
        return out[: M * rows.value].reshape(M, rows.value)

    def set_option(self, key, value):
        B.check(self.lib.ub_set_option(self.handle, key.encode(), int(value)))

    def layout(self):
        out = (C.c_int32 * 80)()
        B.check(self.lib.ub_workspace_layout(self.handle, self.flags, out))
        return dict(zip(LAYOUT_FIELDS, list(out)))

    def workspace_view(self, Bn):
        """Workspace of the last device solve as a [B, total] tensor (debugging / tests)."""
        L = self.layout()
        ws = self._ws.view(self.torch_dtype)
        return ws[: Bn * L["total"]].view(Bn, L["total"]), L

    def workspace_block(self, Bn, name, count):
        """`count` elements of block `name` for each of the Bn static slots of the last device solve, as float64
        numpy [Bn, count]; blocks in LAYOUT_DOUBLE_BLOCKS hold doubles even under the fp32 kernels."""
        import torch
        ws, L = self.workspace_view(Bn)
        if name in LAYOUT_DOUBLE_BLOCKS and self.precision == "f32":
            raw = ws[:, L[name]: L[name] + 2 * count].contiguous()
            return raw.view(torch.float64).cpu().numpy()
        return ws[:, L[name]: L[name] + count].double().cpu().numpy()
