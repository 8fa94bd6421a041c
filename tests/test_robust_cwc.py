"""Contact-wrench-cone verification (upright_robust modelling.py:107-135,
process_sim_runs.py:208-246): the face form produced by facet enumeration against an independent membership test
(non-negative least squares on the span form), and the physics it encodes."""
import copy

import numpy as np
import pytest

import oracle
from upright_b200 import problem_io, robust


def _lin(desc, x):
    nu = oracle.dims(desc)["nu"]
    return oracle.linearize(desc, x, np.zeros(nu))


def _variants():
    d1, meta = problem_io.load_fixture("cfg2_thing_demo")           # frictionless contacts: a flat (rank 3) cone
    d3 = copy.deepcopy(d1)
    d3.nf = 3                                                       # the same contacts with friction pyramids: rank 6
    d5, _ = problem_io.load_fixture("cfg5_thing_robust8")           # eight CoM-vertex bodies, one cone each
    return [("frictionless", d1, meta), ("pyramids", d3, meta), ("robust8", d5, meta)]


@pytest.mark.parametrize("label,desc,meta", _variants(), ids=lambda v: v if isinstance(v, str) else "")
def test_face_form_equals_span_form(label, desc, meta):
    x = np.array(meta["x0"], dtype=float)
    lin = _lin(desc, x)
    H = robust.span_form(desc, lin["Df"])
    A = robust.face_form(desc, lin["Df"])
    A = A / np.linalg.norm(A, axis=1, keepdims=True)
    rng = np.random.default_rng(0)
    n_in = n_out = 0
    for trial in range(200):
        if trial % 2 == 0:      # inside by construction
            w = H @ rng.uniform(0.0, 1.0, H.shape[1])
        else:                   # a generic wrench near the cone
            w = H @ rng.uniform(0.0, 1.0, H.shape[1]) + 0.3 * np.abs(H).max() * rng.standard_normal(H.shape[0])
        dist = robust.in_cone_nnls(H, w)
        viol = float((A @ w).max())
        if dist < 1e-9:
            assert viol < 1e-7, (label, trial, viol, dist)
            n_in += 1
        else:
            # outside: some face is violated, and by no more than the distance to the cone (unit normals)
            assert 1e-10 < viol <= dist + 1e-9, (label, trial, viol, dist)
            n_out += 1
    assert n_in >= 100 and n_out >= 50


def test_wrench_cone_physics():
    """Level tray at rest: the weight is carried (inside the cone).  A base acceleration of 0.05 g is carried by
    friction (mu = 0.234); 0.5 g is not."""
    d1, meta = problem_io.load_fixture("cfg2_thing_demo")
    desc = copy.deepcopy(d1)
    desc.nf = 3
    x = np.array(meta["x0"], dtype=float)
    A = robust.face_form(desc, _lin(desc, x)["Df"])
    A = A / np.linalg.norm(A, axis=1, keepdims=True)
    nq = desc.nq

    def violation(ax):
        xs = x.copy()
        xs[2 * nq] = ax                      # acceleration of the first base joint
        return float((A @ -_lin(desc, xs)["g"]).max())

    assert violation(0.0) < -1e-3
    assert violation(0.05 * 9.81) < 0.0
    assert violation(0.5 * 9.81) > 1e-2
    # frictionless contacts carry no lateral acceleration at all
    A1 = robust.face_form(d1, _lin(d1, x)["Df"])
    A1 = A1 / np.linalg.norm(A1, axis=1, keepdims=True)
    xs = x.copy()
    xs[2 * nq] = 0.05 * 9.81
    assert float((A1 @ -_lin(d1, xs)["g"]).max()) > 1e-2


def test_coupled_arrangement_is_rejected():
    desc, meta = problem_io.load_fixture("cfg3_thing_box_arch")
    lin = _lin(desc, np.array(meta["x0"], dtype=float))
    with pytest.raises(NotImplementedError):
        robust.face_form(desc, lin["Df"])
