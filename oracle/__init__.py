"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle (oracle/oracle.cpp).

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs may import this package.  The product
(`upright_b200`) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_lib = None


def build(force=False):
    so = _DIR / "liboracle.so"
    srcs = [_DIR / n for n in ("oracle.cpp", "model.h", "smallmath.h", "reduced_lab.h")] + [_DIR.parent / "include" / "upright_b200.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["make", "-C", str(_DIR), "-s", "liboracle.so"])
    return so


def lib():
    global _lib
    if _lib is None:
        so = _DIR / "liboracle.so"
        if not so.exists():
            build()
        _lib = C.CDLL(str(so))
        _lib.oracle_qp_dump.restype = C.c_int64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def dims(desc):
    out = (C.c_int32 * 8)()
    lib().oracle_dims(C.byref(desc), out)
    keys = ("nx", "nu", "n_eq", "n_ineq", "n_term", "N", "nb", "nc")
    return dict(zip(keys, out))


def solve_batch(desc, x0, target, body_params=None, X=None, U=None, warm=False, want_gains=False, nthreads=None):
    d = dims(desc)
    x0 = _f64(np.atleast_2d(x0))
    Bn = x0.shape[0]
    target = _f64(np.asarray(target).reshape(Bn, d["N"] + 1, target_stride(desc)))
    body_params = _f64(body_params)
    if X is None or not warm:
        X = np.zeros((Bn, d["N"] + 1, d["nx"]))
        U = np.zeros((Bn, d["N"], d["nu"]))
    else:
        X, U = _f64(X).copy(), _f64(U).copy()
    K = np.zeros((Bn, d["N"], d["nu"], d["nx"])) if want_gains else None
    status = np.zeros(Bn, dtype=np.int32)
    stats = np.zeros((Bn, 8))
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    lib().oracle_solve_batch(C.byref(desc), C.c_int32(Bn), _p(x0), _p(target), _p(body_params), _p(X), _p(U), _p(K),
                             _p(status), _p(stats), C.c_uint32(2 if warm else 0), C.c_int32(nthreads))
    out = dict(X=X, U=U, status=status, stats=stats)
    if want_gains:
        out["K"] = K
    return out


def fk(desc, x):
    out = np.zeros(24 + 3 * desc.n_spheres)
    lib().oracle_fk(C.byref(desc), _p(_f64(x)), _p(out))
    return dict(r=out[0:3], C=out[3:12].reshape(3, 3), v=out[12:15], w=out[15:18], a=out[18:21], alpha=out[21:24],
                spheres=out[24:].reshape(-1, 3))


def linearize(desc, x, u, body_params=None):
    d = dims(desc)
    nq = desc.nq
    nfc = d["nu"] - nq
    nfric = 5 * desc.nc if (desc.balancing_enabled and desc.nf == 3) else 0
    nobs = desc.n_pairs if desc.obstacles_enabled else 0
    o = dict(g=np.zeros(d["n_eq"]), C=np.zeros((d["n_eq"], d["nx"])), Df=np.zeros((d["n_eq"], nfc)), r=np.zeros(3),
             Jp=np.zeros((3, nq)), hfric=np.zeros(nfric), Ffric=np.zeros((nfric, nfc)), hobs=np.zeros(nobs),
             Jobs=np.zeros((nobs, d["nx"])))   # dense over q (+ the position block of a dynamic obstacle)
    lib().oracle_linearize(C.byref(desc), _p(_f64(x)), _p(_f64(u)), _p(_f64(body_params)), _p(o["g"]), _p(o["C"]),
                           _p(o["Df"]), _p(o["r"]), _p(o["Jp"]), _p(o["hfric"]), _p(o["Ffric"]), _p(o["hobs"]),
                           _p(o["Jobs"]))
    o["Jobs_full"] = o["Jobs"]
    o["Jobs"] = o["Jobs"][:, :nq]
    return o


def target_stride(desc):
    """Columns of a target row: 3 (desired position), or 7 (+ desired quaternion x y z w) when the orientation part
    of the end-effector weight is non-zero."""
    return 7 if any(desc.ee_weight[i] != 0 for i in (3, 4, 5)) else 3


def orientation_error(desc, x, qref):
    """End-effector orientation error (ocs2 quaternionDistance of the measured against the desired quaternion
    [x y z w]) and its Jacobian over q."""
    e, J = np.zeros(3), np.zeros((3, desc.nq))
    lib().oracle_orientation_error(C.byref(desc), _p(_f64(x)), _p(_f64(qref)), _p(e), _p(J))
    return e, J


def projectile(desc, x):
    """Projectile-path rows of one knot: values h, Jacobian J over the full state, times of closest approach."""
    d = dims(desc)
    n = desc.n_projectile_links
    o = dict(h=np.zeros(n), J=np.zeros((n, d["nx"])), tclose=np.zeros(n))
    got = lib().oracle_projectile(C.byref(desc), _p(_f64(x)), _p(o["h"]), _p(o["J"]), _p(o["tclose"]))
    return {k: v[:got] for k, v in o.items()}


def performance(desc, target, X, U, body_params=None):
    out = np.zeros(7)
    lib().oracle_performance(C.byref(desc), _p(_f64(target)), _p(_f64(body_params)), _p(_f64(X)), _p(_f64(U)), _p(out))
    return dict(zip(("cost", "dyn_sse", "eq_sse", "ineq_sse", "violation", "max_eq", "min_margin"), out))


def qp_dump(desc, target, X, U, body_params=None):
    """Stage-wise QP data of the first SQP iteration: list of dicts with
    H, g, b, A (rows), c, lb, ub, rho, hard."""
    d = dims(desc)
    N, nx, nu = d["N"], d["nx"], d["nu"]
    rows = (C.c_int32 * (N + 1))()
    args = (C.byref(desc), _p(_f64(target)), _p(_f64(body_params)), _p(_f64(X)), _p(_f64(U)))
    n = lib().oracle_qp_dump(*args, None, rows)
    buf = np.zeros(n)
    lib().oracle_qp_dump(*args, _p(buf), rows)
    out, o = [], 0
    for k in range(N + 1):
        nz = nx + (nu if k < N else 0)
        H = buf[o:o + nz * nz].reshape(nz, nz); o += nz * nz
        g = buf[o:o + nz]; o += nz
        b = buf[o:o + nx]; o += nx
        R = buf[o:o + rows[k] * (nz + 5)].reshape(rows[k], nz + 5); o += rows[k] * (nz + 5)
        out.append(dict(H=H, g=g, b=b, A=R[:, :nz], c=R[:, nz], lb=R[:, nz + 1], ub=R[:, nz + 2], rho=R[:, nz + 3],
                        hard=R[:, nz + 4] > 0))
    return out


def qp_step(desc, target, X, U, body_params=None):
    d = dims(desc)
    dX = np.zeros((d["N"] + 1, d["nx"]))
    dU = np.zeros((d["N"], d["nu"]))
    info = np.zeros(4)
    lib().oracle_qp_step(C.byref(desc), _p(_f64(target)), _p(_f64(body_params)), _p(_f64(X)), _p(_f64(U)), _p(dX),
                         _p(dU), _p(info))
    return dX, dU, dict(iters=int(info[0]), converged=bool(info[1]), decrement=info[2], hard_infeas=info[3])


def qp_step_precision(desc, target, X, U, mode, body_params=None):
    """QP step of the first SQP iteration in a chosen arithmetic (precision study, tools/precision_lab.py):
    mode 0 = fp64, 1 = fp32 throughout (the kernels' arithmetic), 2 = fp32 Newton matrices / factors / direction with
    the iterate, slack records and residuals in fp64."""
    d = dims(desc)
    dX, dU, info = np.zeros((desc.N + 1, d["nx"])), np.zeros((desc.N, d["nu"])), np.zeros(5)
    lib().oracle_qp_step_precision(C.byref(desc), _p(_f64(target)), _p(_f64(body_params)), _p(_f64(X)), _p(_f64(U)),
                                   C.c_int32(int(mode)), _p(dX), _p(dU), _p(info))
    return dict(dX=dX, dU=dU, iters=int(info[0]), converged=bool(info[1]), failed=int(info[2]), mu=info[3], rd=info[4])
