#!/usr/bin/env python3
"""Generate tests/golden/*.json by RUNNING THE REFERENCE'S OWN set-up code.

The reference's Python set-up stage (`upright_core.parsing.parse_control_objects`
and the polyhedron contact search) is importable in the build container once
its absent third-party imports are stubbed: rospkg, xacrodoc,
mobile_manipulation_central (unused by the functions called here),
spatialmath.base (plain rotation helpers, re-implemented below from their
documented behaviour) and the pybind module upright_core.bindings (two plain
structs).  Nothing of the reference is copied: it is imported from
/root/reference, executed, and only its OUTPUTS are stored.

Outputs:
  tests/golden/parsing_config.json      the reference's tests/config.yaml as a dict
  tests/golden/control_objects.json     bodies + contact points per arrangement
  tests/golden/contacts_polyhedron.json contact manifolds of the polyhedron cases
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import yaml

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def install_stubs():
    def q2r(q, order="sxyz"):
        q = np.asarray(q, dtype=float)
        if order == "xyzs":
            x, y, z, s = q
        else:
            s, x, y, z = q
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - s * z), 2 * (x * z + s * y)],
            [2 * (x * y + s * z), 1 - 2 * (x * x + z * z), 2 * (y * z - s * x)],
            [2 * (x * z - s * y), 2 * (y * z + s * x), 1 - 2 * (x * x + y * y)]])

    def r2q(R, order="sxyz"):
        R = np.asarray(R, dtype=float)
        s = 0.5 * np.sqrt(max(0.0, 1.0 + np.trace(R)))
        if s > 1e-8:
            v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (4 * s)
        else:
            i = int(np.argmax(np.diag(R)))
            j, k = (i + 1) % 3, (i + 2) % 3
            v = np.zeros(3)
            v[i] = 0.5 * np.sqrt(max(0.0, 1 + R[i, i] - R[j, j] - R[k, k]))
            v[j] = (R[j, i] + R[i, j]) / (4 * v[i])
            v[k] = (R[k, i] + R[i, k]) / (4 * v[i])
        return np.concatenate((v, [s])) if order == "xyzs" else np.concatenate(([s], v))

    def qunit(q):
        q = np.asarray(q, dtype=float)
        return q / np.linalg.norm(q)

    def rot(axis):
        def f(a):
            c, s = np.cos(a), np.sin(a)
            return {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]),
                    "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
                    "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis].astype(float)
        return f

    sm = types.ModuleType("spatialmath")
    base = types.ModuleType("spatialmath.base")
    base.q2r, base.r2q, base.qunit = q2r, r2q, qunit
    base.rotx, base.roty, base.rotz = rot("x"), rot("y"), rot("z")
    sm.base = base
    sys.modules["spatialmath"], sys.modules["spatialmath.base"] = sm, base

    rospkg = types.ModuleType("rospkg")

    class RosPack:
        def get_path(self, pkg):
            return str(REF / pkg)
    rospkg.RosPack = RosPack
    sys.modules["rospkg"] = rospkg
    xd = types.ModuleType("xacrodoc")
    xd.XacroDoc = object
    sys.modules["xacrodoc"] = xd
    sys.modules["mobile_manipulation_central"] = types.ModuleType("mobile_manipulation_central")
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))

    bindings = types.ModuleType("upright_core.bindings")

    class RigidBody:
        def __init__(self, mass, inertia, com):
            self.mass, self.inertia, self.com = mass, np.array(inertia), np.array(com)

    class ContactPoint:
        pass
    bindings.RigidBody, bindings.ContactPoint = RigidBody, ContactPoint
    sys.modules["upright_core.bindings"] = bindings
    # package shell so that `upright_core/__init__.py` (which imports logging -> matplotlib) is bypassed
    pkg = types.ModuleType("upright_core")
    pkg.__path__ = [str(REF / "upright_core" / "src" / "upright_core")]
    sys.modules["upright_core"] = pkg
    pkg.bindings = bindings


def dump_objects(bodies, contacts):
    return {
        "bodies": {n: {"mass": float(b.mass), "com": np.asarray(b.com).tolist(), "inertia": np.asarray(b.inertia).tolist()}
                   for n, b in bodies.items()},
        "contacts": [{"object1_name": c.object1_name, "object2_name": c.object2_name, "mu": float(c.mu),
                      "r_co_o1": np.asarray(c.r_co_o1).tolist(), "r_co_o2": np.asarray(c.r_co_o2).tolist(),
                      "normal": np.asarray(c.normal).tolist(), "span": np.asarray(c.span).tolist()} for c in contacts],
    }


def main():
    install_stubs()
    import importlib
    parsing = importlib.import_module("upright_core.parsing")
    polyhedron = importlib.import_module("upright_core.polyhedron")
    OUT.mkdir(parents=True, exist_ok=True)

    with open(REF / "upright_core" / "tests" / "config.yaml") as f:
        test_cfg = yaml.safe_load(f)
    with open(OUT / "parsing_config.json", "w") as f:
        json.dump(test_cfg, f, indent=1)

    results = {}
    import copy
    for arr in ("box", "cylinder_box", "wedge_box"):
        cfg = copy.deepcopy(test_cfg)
        cfg["balancing"]["arrangement"] = arr
        results[f"tests/{arr}"] = dump_objects(*parsing.parse_control_objects(cfg))
    for demo, arrs in (("upright_cmd/config/demos/thing_demo.yaml", ["pink_bottle", "box_arch", "foam_die1", "foam_die2"]),
                       ("upright_robust/config/demos/_base.yaml", ["box3_robust"])):
        ctrl = parsing.load_config(str(REF / demo))["controller"]
        for arr in arrs:
            cfg = copy.deepcopy(ctrl)
            cfg["balancing"]["arrangement"] = arr
            results[f"{demo}::{arr}"] = dump_objects(*parsing.parse_control_objects(cfg))
    with open(OUT / "control_objects.json", "w") as f:
        json.dump(results, f, indent=1)

    # polyhedron contact manifolds (cases of upright_core/tests/test_polyhedron.py:182-243)
    P = polyhedron.ConvexPolyhedron
    cases = {}
    b1 = P.box([0.5, 0.5, 0.5])
    b2 = P.box([0.5, 0.5, 0.5], position=None).transform(translation=np.array([0.5, 0.5, 1.0]))
    V, n = polyhedron.axis_aligned_contact(b1, b2)
    cases["box_box_offset"] = {"points": V.tolist(), "normal": n.tolist()}
    w = P.wedge([0.5, 0.5, 0.5])
    C = sys.modules["spatialmath.base"].roty(-np.pi / 4)
    b3 = P.box([0.25, 0.25, 0.25]).transform(rotation=C)
    b3 = b3.transform(translation=-b3.max_vertex_along_axis(np.array([-1.0, 0, -1.0])) * 0 + np.array([0.25 * np.sqrt(2), 0, 0.25 * np.sqrt(2)]))
    V, n = polyhedron.axis_aligned_contact(w, b3)
    cases["wedge_box_slope"] = None if V is None else {"points": V.tolist(), "normal": n.tolist()}
    # the cases of upright_core/tests/test_polyhedron.py:182-243, as written there
    def record(name, a, b):
        V, n = polyhedron.axis_aligned_contact(a, b)
        cases[name] = None if V is None else {"points": np.asarray(V).tolist(), "normal": np.asarray(n).tolist()}

    roty, rotz = sys.modules["spatialmath.base"].roty, sys.modules["spatialmath.base"].rotz
    box1 = P.box([1, 1, 1])
    box2 = P.box([0.5, 0.5, 0.5]).transform(translation=[0.5, 0.5, 1.5])
    record("test_box_box_contact", box1, box2)
    record("test_box_box_contact/penetrating", box1, box2.transform(translation=[0, 0, -0.1]))
    record("test_box_box_contact/separated", box1, box2.transform(translation=[0, 0, 0.1]))
    nrm = np.array([1.0, 0, 1.0]) / np.sqrt(2.0)
    record("test_wedge_box_contact", P.wedge([1, 1, 1]), P.box([1, 1, 1]).transform(rotation=roty(-np.pi / 4), translation=nrm))
    thin = P.box([0.03, 0.03, 0.3]).transform(rotation=rotz(np.pi / 4))
    dx = thin.distance_from_centroid_to_boundary([1, 0, 0])
    record("test_line_contact", thin, P.box([0.1, 0.1, 0.1]).transform(translation=[dx + 0.1, 0, 0]))
    cases["test_line_contact"]["dx"] = float(dx)
    with open(OUT / "contacts_polyhedron.json", "w") as f:
        json.dump(cases, f, indent=1)
    print("golden files written to", OUT)
    for k, v in results.items():
        print(f"  {k}: {len(v['bodies'])} bodies, {len(v['contacts'])} contacts")


if __name__ == "__main__":
    main()
