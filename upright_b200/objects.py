"""Arrangement -> rigid bodies + contact points (set-up stage of the path).

Restates `parse_control_objects` and helpers of
`upright_core/src/upright_core/parsing.py:162-410`: objects are stacked on
their parent's top face, cylinders become inscribed boxes turned 45° about z
(`:227-249`), wedges carry their centroid offset (`:311-316`), each pairwise
contact yields the clipped contact polygon's vertices as contact points with
the friction margin subtracted and the support area inset applied in the
tangent plane (`:162-220`).  Frames: everything is expressed in the
end-effector (tray) frame, as `contact_constraints.h:122` assumes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import geometry as geo
from .config import parse_support_offset


@dataclass
class RigidBody:
    """Mirror of `upright::RigidBody` (upright_core/include/upright_core/rigid_body.h:29-66)."""

    mass: float
    inertia: np.ndarray
    com: np.ndarray

    def get_parameters(self) -> np.ndarray:
        """[m, m*com, vech(I)] (rigid_body.h:50-54)."""
        I = self.inertia
        return np.concatenate(
            ([self.mass], self.mass * self.com, [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]])
        )

    @classmethod
    def from_parameters(cls, p):
        p = np.asarray(p, dtype=float)
        I = np.array([[p[4], p[5], p[6]], [p[5], p[7], p[8]], [p[6], p[8], p[9]]])
        return cls(float(p[0]), I, p[1:4] / p[0])


@dataclass
class ContactPoint:
    """Mirror of `upright::ContactPoint` (upright_core/include/upright_core/contact.h:10-48)."""

    object1_name: str = ""
    object2_name: str = ""
    mu: float = 0.0
    r_co_o1: np.ndarray = field(default_factory=lambda: np.zeros(3))
    r_co_o2: np.ndarray = field(default_factory=lambda: np.zeros(3))
    normal: np.ndarray = field(default_factory=lambda: np.zeros(3))
    span: np.ndarray = field(default_factory=lambda: np.zeros((2, 3)))


@dataclass
class _Placed:
    body: RigidBody
    box: geo.ConvexPolyhedron
    parent: str | None
    fixture: bool


def _normalise_shape_keys(type_confs):
    """Old config format nests shape parameters under `shape: {type: ...}`."""
    for conf in type_confs.values():
        shape = conf["shape"]
        if isinstance(shape, dict):
            shape = dict(shape)
            conf["shape"] = shape.pop("type")
            conf.update(shape)


def local_half_extents(conf):
    shape = conf["shape"].lower()
    if shape in ("cuboid", "wedge"):
        return 0.5 * np.asarray(conf["side_lengths"], dtype=float)
    if shape == "cylinder":
        w = np.sqrt(2.0) * conf["radius"]
        return 0.5 * np.array([w, w, conf["height"]])
    raise ValueError(f"unsupported shape {shape}")


def make_box(conf, position=None, rotation=None):
    R = np.eye(3) if rotation is None else rotation
    shape = conf["shape"].lower()
    he = local_half_extents(conf)
    if shape == "wedge":
        poly = geo.ConvexPolyhedron.wedge(he)
    else:
        poly = geo.ConvexPolyhedron.box(he)
        if shape == "cylinder":
            R = R @ geo.rotz(np.pi / 4)
    return poly.transform(translation=position, rotation=R)


def shape_inertia(mass, conf):
    shape = conf["shape"].lower()
    if shape == "cylinder":
        return geo.cylinder_inertia(mass, conf["radius"], conf["height"])
    if shape == "cuboid":
        return geo.cuboid_inertia(mass, conf["side_lengths"])
    if shape == "wedge":
        return geo.wedge_inertia(mass, conf["side_lengths"])
    raise ValueError(f"unsupported shape {shape}")


def _place_object(conf, base_position, quat):
    mass = conf["mass"]
    C = geo.quat_to_rot(quat)
    local_com = np.array(conf["com_offset"], dtype=float)
    if conf["shape"].lower() == "wedge":
        hx, _, hz = 0.5 * np.asarray(conf["side_lengths"], dtype=float)
        local_com = local_com + np.array([-hx, 0.0, -hz]) / 3.0
    if "inertia" in conf:
        I_local = np.asarray(conf["inertia"], dtype=float)
        if I_local.shape == (3,):
            I_local = np.diag(I_local)
        elif I_local.shape != (3, 3):
            raise ValueError(f"inertia has wrong shape {I_local.shape}")
    elif "inertia_diag" in conf:
        I_local = np.diag(conf["inertia_diag"])
    else:
        I_local = shape_inertia(mass, conf)
    inertia = C @ I_local @ C.T

    dz = make_box(conf, rotation=C).distance_from_centroid_to_boundary(np.array([0.0, 0.0, -1.0]))
    reference = np.asarray(base_position, dtype=float) + np.array([0.0, 0.0, dz])
    body = RigidBody(mass, inertia, reference + C @ local_com)
    return body, make_box(conf, reference, C)


def _contact_points(placed, contact_conf):
    points = []
    for spec in contact_conf:
        n1, n2 = spec["first"], spec["second"]
        mu = spec["mu"] - spec.get("mu_margin", 0)
        inset = spec.get("support_area_inset", 0)
        box1, box2 = placed[n1].box, placed[n2].box
        verts, normal = geo.axis_aligned_contact(box1, box2, tol=1e-7)
        assert verts is not None, f"no contact found between {n1} and {n2}"
        span = geo.plane_span(normal)
        for r in verts:
            if placed[n1].fixture:
                r1 = r.copy()
            else:
                t = span @ (r - box1.position)
                r1 = r + (geo.inset_vertex(t, inset) - t) @ span
            t = span @ (r - box2.position)
            r2 = r + (geo.inset_vertex(t, inset) - t) @ span
            points.append(ContactPoint(n1, n2, float(mu), r1, r2, np.array(normal), span.copy()))
    return points


def parse_control_objects(ctrl_conf):
    """-> (bodies: dict name->RigidBody (non-fixtures), contacts: list[ContactPoint])."""
    arrangement = ctrl_conf["arrangements"][ctrl_conf["balancing"]["arrangement"]]
    type_confs = ctrl_conf["objects"]
    _normalise_shape_keys(type_confs)

    ee_conf = type_confs["ee"]
    ee_box = make_box(ee_conf, np.array(ee_conf["position"], dtype=float))
    placed = {"ee": _Placed(RigidBody(1.0, np.eye(3), ee_box.position), ee_box, None, True)}

    up = np.array([0.0, 0.0, 1.0])
    for inst in arrangement["objects"]:
        name = inst["name"]
        if name in placed:
            raise ValueError(f"multiple control objects named {name}")
        conf = type_confs[inst["type"]]
        quat = np.asarray(inst.get("orientation", [0, 0, 0, 1]), dtype=float)
        quat = quat / np.linalg.norm(quat)
        parent_box = placed[inst["parent"]].box
        base = parent_box.position.copy()
        if "offset" in inst:
            base[:2] += parse_support_offset(inst["offset"])
        base[2] += parent_box.distance_from_centroid_to_boundary(up)
        body, box = _place_object(conf, base, quat)
        placed[name] = _Placed(body, box, inst["parent"], bool(inst.get("fixture", False)))

    contacts = _contact_points(placed, arrangement.get("contacts", []))
    bodies = {n: p.body for n, p in placed.items() if not p.fixture}
    return bodies, contacts
