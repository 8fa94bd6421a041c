"""Set-up stage against the reference: (i) the known answers of the
reference's own tests (upright_core/tests/test_parsing.py:23-181, test_math.py)
and (ii) golden outputs produced by running the reference's
parse_control_objects in the build container (tools/gen_golden.py)."""
import copy
import json
from pathlib import Path

import numpy as np
import pytest

from upright_b200 import config as cfg
from upright_b200 import geometry as geo
from upright_b200 import objects

GOLD = Path(__file__).parent / "golden"


def unordered_close(A, B, tol=1e-9):
    A, B = np.asarray(A), np.asarray(B)
    if A.shape != B.shape:
        return False
    used = np.zeros(len(B), dtype=bool)
    for a in A:
        d = np.linalg.norm(B - a, axis=1)
        d[used] = np.inf
        i = int(np.argmin(d))
        if d[i] > tol:
            return False
        used[i] = True
    return True


def test_number_and_array_dsl():
    assert cfg.parse_number("1e-2") == 1e-2
    assert cfg.parse_number("2pi") == 2 * np.pi
    assert np.isclose(cfg.parse_number(1), 1)
    assert np.allclose(cfg.parse_array([1, 2, 3]), [1, 2, 3])
    assert np.allclose(cfg.parse_array(["1rep3"]), [1, 1, 1])
    assert np.allclose(cfg.parse_array(["1rep3", "1pi"]), [1, 1, 1, np.pi])
    assert np.allclose(cfg.parse_diag_matrix_dict({"scale": 2, "diag": ["1rep3"]}), 2 * np.eye(3))


def test_support_offset():
    th = 0.25 * np.pi
    assert np.allclose(cfg.parse_support_offset({"x": 1, "y": 1}), [1, 1])
    assert np.allclose(cfg.parse_support_offset({"r": 1, "θ": th}), [np.cos(th), np.sin(th)])
    assert np.allclose(cfg.parse_support_offset({"x": 1, "y": 1, "r": 1, "θ": th}), [np.cos(th) + 1, np.sin(th) + 1])
    with pytest.raises(ValueError):
        cfg.parse_support_offset({"r": 1})
    with pytest.raises(ValueError):
        cfg.parse_support_offset({"θ": 1})


def test_math_helpers():
    assert np.allclose(geo.skew3([1, 2, 3]), [[0, -3, 2], [3, 0, -1], [-2, 1, 0]])
    assert np.allclose(geo.quat_to_rot([0, 0, 0, 1]), np.eye(3))
    assert np.allclose(geo.rot_to_quat(np.eye(3)), [0, 0, 0, 1])
    q = np.array([1.0, 2, 3, 4]) / np.linalg.norm([1, 2, 3, 4])
    assert np.allclose(geo.rot_to_quat(geo.quat_to_rot(q)), q)
    assert np.isclose(geo.quat_angle(np.array([0, 0, 0, 1.0])), 0)
    assert np.isclose(geo.quat_angle(np.array([1, 0, 0, 1.0]) / np.sqrt(2)), 0.5 * np.pi)
    v = np.array([1.0, 1.0])
    assert np.allclose(geo.inset_vertex(v, 0.5 * np.sqrt(2)), [0.5, 0.5])
    with pytest.raises(ValueError):
        geo.inset_vertex(v, 2.0)
    assert np.allclose(geo.inset_vertex_abs(np.array([1.0, -2.0]), 0.5), [0.5, -1.5])
    S = geo.plane_span(np.array([0.0, 0, 1]))
    assert S.shape == (2, 3) and np.allclose(S @ [0, 0, 1], 0) and np.allclose(S @ S.T, np.eye(2))


def load_test_config():
    with open(GOLD / "parsing_config.json") as f:
        return json.load(f)


def test_box_known_answers():
    c = load_test_config()
    c["balancing"]["arrangement"] = "box"
    bodies, contacts = objects.parse_control_objects(c)
    box = bodies["box"]
    assert np.isclose(box.mass, 1.0)
    assert np.allclose(box.com, [0, 0, 0.1])
    assert np.allclose(box.inertia, geo.cuboid_inertia(1.0, [0.2, 0.2, 0.2]))
    assert len(contacts) == 4
    for ct in contacts:
        assert np.allclose(ct.normal, [0, 0, -1])
        assert np.allclose(ct.span @ ct.normal, [0, 0])
        assert np.isclose(ct.mu, 0.45)
        assert ct.object1_name == "ee" and ct.object2_name == "box"
    expected = np.array([[0.1, 0.1, 0], [0.1, -0.1, 0], [-0.1, -0.1, 0], [-0.1, 0.1, 0]])
    assert unordered_close([ct.r_co_o1 for ct in contacts], expected)
    assert unordered_close([ct.r_co_o2 for ct in contacts], expected)


def points_by_name(names, contacts):
    pts = {n: [] for n in names}
    for c in contacts:
        pts[c.object1_name].append(c.r_co_o1)
        pts[c.object2_name].append(c.r_co_o2)
    return {n: np.array(v) for n, v in pts.items()}


def test_cylinder_box_known_answers():
    c = load_test_config()
    c["balancing"]["arrangement"] = "cylinder_box"
    _, contacts = objects.parse_control_objects(c)
    assert len(contacts) == 10
    for ct in contacts:
        if ct.object1_name == "ee":
            assert np.allclose(ct.normal, [0, 0, -1])
        else:
            assert np.allclose(np.abs(ct.normal), [1, 0, 0])
    pts = points_by_name(["ee", "box", "cylinder"], contacts)
    ee = [[0, 0, 0], [-0.03, 0.03, 0], [-0.06, 0, 0], [-0.03, -0.03, 0], [0, -0.1, 0], [0.2, -0.1, 0], [0.2, 0.1, 0], [0, 0.1, 0]]
    box = [[0, -0.1, 0], [0.2, -0.1, 0], [0.2, 0.1, 0], [0, 0.1, 0], [0, 0, 0], [0, 0, 0.2]]
    cyl = [[0, 0, 0], [-0.03, 0.03, 0], [-0.06, 0, 0], [-0.03, -0.03, 0], [0, 0, 0], [0, 0, 0.2]]
    assert unordered_close(pts["ee"], ee) and unordered_close(pts["box"], box) and unordered_close(pts["cylinder"], cyl)


def test_wedge_box_known_answers():
    c = load_test_config()
    c["balancing"]["arrangement"] = "wedge_box"
    bodies, contacts = objects.parse_control_objects(c)
    assert np.allclose(bodies["wedge"].com, [-0.05, 0, 0.1])
    assert len(contacts) == 8
    z = np.array([0.0, 0, 1])
    Cz = geo.roty(np.pi / 4) @ z
    for ct in contacts:
        assert np.allclose(ct.normal, -z if ct.object1_name == "ee" else -Cz)
    pts = points_by_name(["ee", "wedge", "box"], contacts)
    a = np.sqrt(0.02)
    wedge = [[0.15, 0.15, 0], [0.15, -0.15, 0], [-0.15, -0.15, 0], [-0.15, 0.15, 0], [0, 0.1, 0.15], [0, -0.1, 0.15],
             [-a, 0.1, 0.15 + a], [-a, -0.1, 0.15 + a]]
    box = [[0, 0.1, 0.15], [0, -0.1, 0.15], [-a, -0.1, 0.15 + a], [-a, 0.1, 0.15 + a]]
    assert unordered_close(pts["wedge"], wedge) and unordered_close(pts["box"], box)


def _check_against_reference(ctrl, key, gold):
    bodies, contacts = objects.parse_control_objects(ctrl)
    g = gold[key]
    assert sorted(bodies) == sorted(g["bodies"])
    for name, gb in g["bodies"].items():
        assert np.isclose(bodies[name].mass, gb["mass"])
        assert np.allclose(bodies[name].com, gb["com"], atol=1e-12)
        assert np.allclose(bodies[name].inertia, gb["inertia"], atol=1e-12)
    assert len(contacts) == len(g["contacts"])
    # same order, same tangent basis: the force-variable layout depends on both
    for ct, gc in zip(contacts, g["contacts"]):
        assert (ct.object1_name, ct.object2_name) == (gc["object1_name"], gc["object2_name"])
        assert np.isclose(ct.mu, gc["mu"])
        assert np.allclose(ct.r_co_o1, gc["r_co_o1"], atol=1e-12)
        assert np.allclose(ct.r_co_o2, gc["r_co_o2"], atol=1e-12)
        assert np.allclose(ct.normal, gc["normal"], atol=1e-12)
        assert np.allclose(ct.span, gc["span"], atol=1e-12)


@pytest.mark.parametrize("arr", ["box", "cylinder_box", "wedge_box"])
def test_reference_outputs_test_config(arr):
    with open(GOLD / "control_objects.json") as f:
        gold = json.load(f)
    c = load_test_config()
    c["balancing"]["arrangement"] = arr
    _check_against_reference(c, f"tests/{arr}", gold)


@pytest.mark.parametrize("demo,arr", [
    ("upright_cmd/config/demos/thing_demo.yaml", "pink_bottle"),
    ("upright_cmd/config/demos/thing_demo.yaml", "box_arch"),
    ("upright_cmd/config/demos/thing_demo.yaml", "foam_die1"),
    ("upright_cmd/config/demos/thing_demo.yaml", "foam_die2"),
    ("upright_robust/config/demos/_base.yaml", "box3_robust"),
])
def test_reference_outputs_shipped_arrangements(demo, arr):
    """Needs the reference YAML tree (build container only)."""
    ref = Path("/root/reference") / demo
    if not ref.exists():
        pytest.skip("reference tree not present")
    with open(GOLD / "control_objects.json") as f:
        gold = json.load(f)
    ctrl = copy.deepcopy(cfg.load_config(ref)["controller"])
    ctrl["balancing"]["arrangement"] = arr
    _check_against_reference(ctrl, f"{demo}::{arr}", gold)


def test_fixture_contacts_match_reference_outputs():
    """The committed problem fixtures carry exactly the reference's bodies/contacts."""
    from upright_b200 import problem_io
    with open(GOLD / "control_objects.json") as f:
        gold = json.load(f)
    for fix, key in (("cfg2_thing_demo", "upright_cmd/config/demos/thing_demo.yaml::pink_bottle"),
                     ("cfg3_thing_box_arch", "upright_cmd/config/demos/thing_demo.yaml::box_arch"),
                     ("cfg5_thing_robust8", "upright_robust/config/demos/_base.yaml::box3_robust")):
        desc, meta = problem_io.load_fixture(fix)
        g = gold[key]
        names = sorted(g["bodies"])
        assert meta["body_names"] == names
        for b, n in enumerate(names):
            gb = g["bodies"][n]
            p = [desc.body_params[b][j] for j in range(10)]
            I = np.array(gb["inertia"])
            exp = [gb["mass"], *(gb["mass"] * np.array(gb["com"])), I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]
            assert np.allclose(p, exp, atol=1e-12)
        assert desc.nc == len(g["contacts"])
        for i, gc in enumerate(g["contacts"]):
            c = desc.contacts[i]
            assert c.body1 == (names.index(gc["object1_name"]) if gc["object1_name"] in names else -1)
            assert c.body2 == names.index(gc["object2_name"])
            assert np.isclose(c.mu, gc["mu"])
            assert np.allclose(list(c.r_co_o1), gc["r_co_o1"], atol=1e-12)
            assert np.allclose(list(c.r_co_o2), gc["r_co_o2"], atol=1e-12)
            assert np.allclose(list(c.span), np.array(gc["span"]).ravel(), atol=1e-12)


def _polyhedron_cases():
    """The contact cases of upright_core/tests/test_polyhedron.py:182-243, built with this repo's geometry."""
    P = geo.ConvexPolyhedron
    box1 = P.box([1, 1, 1])
    box2 = P.box([0.5, 0.5, 0.5]).transform(translation=[0.5, 0.5, 1.5])
    n = np.array([1.0, 0, 1.0]) / np.sqrt(2.0)
    thin = P.box([0.03, 0.03, 0.3]).transform(rotation=geo.rotz(np.pi / 4))
    dx = thin.distance_from_centroid_to_boundary(np.array([1.0, 0, 0]))
    return {
        "test_box_box_contact": (box1, box2),
        "test_box_box_contact/penetrating": (box1, box2.transform(translation=[0, 0, -0.1])),
        "test_box_box_contact/separated": (box1, box2.transform(translation=[0, 0, 0.1])),
        "test_wedge_box_contact": (P.wedge([1, 1, 1]), P.box([1, 1, 1]).transform(rotation=geo.roty(-np.pi / 4), translation=n)),
        "test_line_contact": (thin, P.box([0.1, 0.1, 0.1]).transform(translation=[dx + 0.1, 0, 0])),
        "box_box_offset": (P.box([0.5, 0.5, 0.5]), P.box([0.5, 0.5, 0.5]).transform(translation=[0.5, 0.5, 1.0])),
    }, dx


def test_polyhedron_contact_manifolds_against_the_reference():
    """Contact points and normals of every case of the reference's own polyhedron tests: the golden file holds what
    `upright_core.polyhedron.axis_aligned_contact` itself returns for them (tools/gen_golden.py), and the known
    answers written in test_polyhedron.py:182-243 are asserted on top."""
    gold = json.loads((GOLD / "contacts_polyhedron.json").read_text())
    cases, dx = _polyhedron_cases()
    for name, (a, b) in cases.items():
        V, n = geo.axis_aligned_contact(a, b)
        want = gold[name]
        if want is None:
            assert V is None and n is None, name
            continue
        assert V is not None, name
        assert np.allclose(n, want["normal"], atol=1e-12), name
        assert unordered_close(V, want["points"]), name
    assert gold["test_line_contact"]["dx"] == pytest.approx(dx, abs=1e-12)
    # the answers the reference's tests state
    V, n = geo.axis_aligned_contact(*cases["test_box_box_contact"])
    assert np.allclose(n, [0, 0, -1]) and unordered_close(V, [[0, 0, 1], [0, 1, 1], [1, 1, 1], [1, 0, 1]])
    V, n = geo.axis_aligned_contact(*cases["test_wedge_box_contact"])
    a = np.sqrt(2) / 2
    assert np.allclose(n, -np.array([a, 0, a])) and unordered_close(V, [[-a, 1, a], [-a, -1, a], [a, -1, -a], [a, 1, -a]])
    V, n = geo.axis_aligned_contact(*cases["test_line_contact"])
    assert np.allclose(n, [-1, 0, 0]) and unordered_close(V, [[dx, 0, -0.1], [dx, 0, 0.1]])
