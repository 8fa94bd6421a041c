"""Contact-wrench-cone verification of planned trajectories.

The reference verifies, off line, that the gravito-inertial wrench of every object along a planned trajectory can be
carried by contact forces inside the friction cones: it builds the SPAN form of the contact wrench cone of an
arrangement (upright_robust/src/upright_robust/modelling.py:107-124 `compute_cwc_span_form`), converts it to FACE form
`A w <= 0` (`compute_cwc_face_form`, :127-135, through `utils.cone_span_to_face_form` = rigeo's `SpanForm.to_face_form`,
i.e. the double-description method of pycddlib — both un-vendored third-party packages, no version pinned in the
reference) and evaluates `max_i A_i w` at every planned state (upright_robust/scripts/process_sim_runs.py:208-246,
`exact_params` branch; with uncertain parameters the same rows are maximised over the parameter polytope with MOSEK,
which for a polytope given by its vertices is the maximum over the vertices — the eight CoM-vertex bodies of
`upright_robust/config/arrangements.yaml`).

Here the pieces map onto what the solver already has:

* the wrench space is the row space of the object-dynamics equalities `g(x, f) = g0(x) + Df f = 0` of the hot path
  (contact_constraints.h:79-194): `-g0(x)` is the (scaled) gravito-inertial wrench the contacts have to supply, the
  columns of `Df` are the wrenches of unit contact forces, both evaluated on the GPU by the probes of `ub_eval`;
* span form: `H = Df blockdiag(S_c)` with `S_c` the extreme rays of the linearised friction pyramid of contact c
  (contact_constraints.h:49-77: `n'f >= 0`, `mu n'f +- s0'f +- s1'f >= 0`, rays `n +- mu s0`, `n +- mu s1`; a
  frictionless contact has the single ray `n`);
* face form: `cone_span_to_face_form` below — facets of a pointed cone by enumeration of (r - 1)-subsets of the rays
  in the r-dimensional span, plus the equalities of its orthogonal complement.  Exact for the per-body cones used here
  (r <= 6, up to a few dozen rays); arrangements whose bodies share contacts have ONE cone of dimension 6 nb for which
  enumeration is hopeless (the reference's cdd call scales the same way) — rejected;
* the check itself is one GEMM over all planned knots of the batch, `[B (N + 1), 6 nb] x [6 nb, faces]`, on the GPU
  (torch.matmul: a plain library GEMM), with the wrenches produced by the probe kernel — no LP, no MOSEK.
"""
from __future__ import annotations

import itertools

import numpy as np


def friction_rays(desc, c):
    """Extreme rays [n_rays, nf] of the admissible forces of contact c in the force coordinates of the input vector:
    nf = 1 (frictionless: magnitude along the normal) -> [[1]]; nf = 3 -> n +- mu s0, n +- mu s1."""
    if desc.nf == 1:
        return np.ones((1, 1))
    ct = desc.contacts[c]
    n = np.array(ct.normal[:3], dtype=float)
    s0, s1 = np.array(ct.span[:3], dtype=float), np.array(ct.span[3:6], dtype=float)
    mu = float(ct.mu)
    return np.stack((n + mu * s0, n - mu * s0, n + mu * s1, n - mu * s1))


def span_form(desc, Df):
    """H [6 nb, n_rays_total]: wrenches (in the row space of the object-dynamics equalities) of the extreme rays of all
    contacts; Df [6 nb, nf nc] is the force block of the constraint Jacobian (probe "object_dynamics_jacobian")."""
    nf, cols = desc.nf, []
    for c in range(desc.nc):
        S = friction_rays(desc, c)                       # [rays, nf]
        cols.append(Df[:, nf * c: nf * (c + 1)] @ S.T)    # [6 nb, rays]
    return np.hstack(cols)


def cone_span_to_face_form(S, tol=1e-9):
    """Face form A (A w <= 0) of the pointed polyhedral cone { S z | z >= 0 }, S [d, n].

    The cone is full-dimensional in the r-dimensional span of its rays (r = rank S): every facet there is spanned by
    r - 1 linearly independent rays with all other rays on one side.  Returns the facet normals mapped back to the
    ambient space followed by +- the normals of the orthogonal complement (the equalities of a flat cone)."""
    S = np.asarray(S, dtype=float)
    d, n = S.shape
    scale = np.linalg.norm(S, axis=0)
    keep = scale > tol * max(1.0, scale.max())
    Sn = S[:, keep] / scale[keep]
    U, sv, _ = np.linalg.svd(Sn, full_matrices=True)
    r = int((sv > 1e-9 * sv[0]).sum())
    Q, Qperp = U[:, :r], U[:, r:]
    Y = Q.T @ Sn                                          # rays in span coordinates [r, n]
    normals = []
    if r == 1:
        normals = [np.array([-np.sign(Y[0, 0])])]         # a half-line: -y <= 0 along the ray, equalities elsewhere
    else:
        m = Y.shape[1]
        if m > 40 or r > 6:
            raise NotImplementedError("cone too large for facet enumeration (bodies that share contacts form one cone)")
        for idx in itertools.combinations(range(m), r - 1):
            sub = Y[:, idx]
            _, s2, vt = np.linalg.svd(sub.T, full_matrices=True)
            if s2.size < r - 1 or s2[r - 2] < 1e-9:
                continue                                  # the subset does not span a hyperplane
            a = vt[-1]                                    # normal of the hyperplane through the subset (and the origin)
            side = a @ Y
            if np.all(side <= 1e-9):
                pass
            elif np.all(side >= -1e-9):
                a = -a
            else:
                continue
            if not any(np.allclose(a, b, atol=1e-7) for b in normals):
                normals.append(a)
    rows = [Q @ a for a in normals]
    for j in range(Qperp.shape[1]):
        rows.append(Qperp[:, j])
        rows.append(-Qperp[:, j])
    return np.array(rows).reshape(-1, d)


def face_form(desc, Df):
    """Face form of the contact wrench cone of the arrangement, body by body (block diagonal).  Bodies that share
    contacts are one coupled cone and are rejected (see the module docstring)."""
    nb, nf = desc.nb, desc.nf
    owner = {}
    for c in range(desc.nc):
        ct = desc.contacts[c]
        bodies = [b for b in (ct.body1, ct.body2) if b >= 0]
        if len(bodies) > 1:
            raise NotImplementedError("bodies that share contacts form one 6 nb-dimensional cone: not enumerated")
        owner.setdefault(bodies[0], []).append(c)
    blocks = []
    for b in range(nb):
        cols = []
        for c in owner.get(b, []):
            cols.append(Df[6 * b: 6 * b + 6, nf * c: nf * (c + 1)] @ friction_rays(desc, c).T)
        Ab = cone_span_to_face_form(np.hstack(cols))
        full = np.zeros((Ab.shape[0], 6 * nb))
        full[:, 6 * b: 6 * b + 6] = Ab
        blocks.append(full)
    return np.vstack(blocks)


def in_cone_nnls(S, w):
    """Independent membership test (test infrastructure): distance of w from { S z | z >= 0 } by NNLS."""
    from scipy.optimize import nnls
    return nnls(np.asarray(S, dtype=float), np.asarray(w, dtype=float))[1]


class WrenchConeVerifier:
    """`max_i A_i w` of the contact wrench cone at every knot of a batch of planned trajectories, on the GPU.

    mpc: `BatchedMPC` (supplies the probes); body parameters are the problem's nominal ones (the reference verifies an
    arrangement with its nominal / vertex parameters, process_sim_runs.py:224-230)."""

    def __init__(self, mpc, desc):
        import torch
        self.mpc, self.desc = mpc, desc
        x = np.zeros((1, mpc.nx))
        G = mpc.eval("object_dynamics_jacobian", x, np.zeros((1, mpc.nu))).reshape(mpc.n_eq, mpc.nx + mpc.nu)
        self.Df = G[:, mpc.nx + desc.nq:]
        self.A = face_form(desc, self.Df)
        nrm = np.linalg.norm(self.A, axis=1, keepdims=True)
        self.A = self.A / np.where(nrm > 0, nrm, 1.0)      # unit normals: violations are distances in wrench space
        self.A_dev = torch.tensor(self.A, dtype=torch.float64, device="cuda")

    def wrenches(self, X):
        """-g0(x) [B, K, 6 nb]: the wrench the contacts have to supply at every given state (forces set to zero)."""
        X = np.asarray(X, dtype=float)
        B, K = X.shape[0], X.shape[1]
        g0 = self.mpc.eval("object_dynamics", X.reshape(B * K, -1)[:, : self.mpc.nx], np.zeros((B * K, self.mpc.nu)))
        return -g0.reshape(B, K, -1)

    def violation(self, X):
        """[B, K] largest face violation per knot (<= 0: the wrench can be carried inside the friction cones)."""
        import torch
        W = torch.tensor(self.wrenches(X), dtype=torch.float64, device="cuda")
        V = torch.matmul(W, self.A_dev.T)                  # [B, K, faces]
        return V.max(dim=2).values.cpu().numpy()
