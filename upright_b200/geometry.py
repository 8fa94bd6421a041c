"""Rigid-body / convex-polyhedron helpers used once at set-up time to turn an
object arrangement into rigid bodies and contact points.

Behavioural restatement of `upright_core/src/upright_core/math.py` (quaternion
xyzw convention `:6`, inertia formulas `:104-141`, `inset_vertex :144-154`,
`plane_span :163-178`) and `upright_core/src/upright_core/polyhedron.py`
(box/wedge construction `:44-93`, separating-axis contact search and polygon
clipping `:446-514`).  Written independently; the axis enumeration order of
the contact search follows the reference because it decides which tangent
basis (`plane_span`) and vertex winding the contact points come out in.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import null_space
from scipy.optimize import linprog

TOL = 1e-8


# ----------------------------------------------------------------- rotations
def rotx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=float)


def roty(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=float)


def rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=float)


def rpy_to_rot(rpy):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = rpy
    return rotz(y) @ roty(p) @ rotx(r)


def skew3(v):
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=float)


def quat_to_rot(q):
    """Unit quaternion [x, y, z, w] -> rotation matrix (math.py:6,59-61)."""
    x, y, z, w = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def rot_to_quat(C):
    """Rotation matrix -> [x, y, z, w] with w >= 0 (math.py:64-66)."""
    C = np.asarray(C, dtype=float)
    tr = np.trace(C)
    if tr > 0:
        s = 2.0 * np.sqrt(1.0 + tr)
        q = [(C[2, 1] - C[1, 2]) / s, (C[0, 2] - C[2, 0]) / s, (C[1, 0] - C[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(C)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = 2.0 * np.sqrt(1.0 + C[i, i] - C[j, j] - C[k, k])
        q = [0.0, 0.0, 0.0, (C[k, j] - C[j, k]) / s]
        q[i] = 0.25 * s
        q[j] = (C[j, i] + C[i, j]) / s
        q[k] = (C[k, i] + C[i, k]) / s
    q = np.array(q)
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_multiply(q0, q1):
    """Hamilton product through rotation matrices, as math.py:69-76 does."""
    return rot_to_quat(quat_to_rot(q0) @ quat_to_rot(q1))


def quat_angle(q):
    return 2 * np.arctan2(np.linalg.norm(q[:3]), q[3])


# ------------------------------------------------------------------ inertias
def cylinder_inertia(mass, radius, height):
    xx = mass * (3 * radius**2 + height**2) / 12.0
    return np.diag([xx, xx, 0.5 * mass * radius**2])


def cuboid_inertia(mass, side_lengths):
    lx, ly, lz = side_lengths
    return mass * np.diag([ly**2 + lz**2, lx**2 + lz**2, lx**2 + ly**2]) / 12.0


def wedge_inertia(mass, side_lengths):
    """Right-triangular prism about its centroid (math.py:122-141)."""
    hx, hy, hz = 0.5 * np.asarray(side_lengths, dtype=float)
    J = np.array(
        [
            [hy**2 / 3 + 2 * hz**2 / 9, 0, hx * hz / 9],
            [0, 2 * hx**2 / 9 + 2 * hz**2 / 9, 0],
            [hx * hz / 9, 0, 2 * hx**2 / 9 + hy**2 / 3],
        ]
    )
    return mass * J


def inset_vertex(v, inset):
    """Pull `v` toward the origin by `inset` (math.py:144-154)."""
    d = np.linalg.norm(v)
    if d <= inset:
        raise ValueError(f"inset {inset} too large for the support area")
    return (d - inset) * v / d


def inset_vertex_abs(v, inset):
    v = np.asarray(v, dtype=float)
    if (np.abs(v) <= inset).any():
        raise ValueError(f"inset {inset} too large for the support area")
    return v - np.sign(v) * inset


def plane_span(normal):
    """(n-1, n) orthonormal rows spanning the plane orthogonal to `normal`.

    The reference takes scipy's SVD null-space basis (math.py:163-178); the
    basis choice fixes the orientation of the linearised friction pyramid, so
    the same call is used here and the result is recorded in the fixtures.
    """
    return null_space(np.asarray(normal, dtype=float)[None, :]).T


# ---------------------------------------------------------------- polyhedra
class ConvexPolyhedron:
    """Convex polyhedron given by vertices and outward face normals."""

    def __init__(self, vertices, normals, position=None, rotation=None):
        self.vertices = np.asarray(vertices, dtype=float)
        self.normals = np.asarray(normals, dtype=float)
        self.position = np.zeros(3) if position is None else np.asarray(position, dtype=float)
        self.rotation = np.eye(3) if rotation is None else np.asarray(rotation, dtype=float)

    @classmethod
    def box(cls, half_extents):
        x, y, z = np.asarray(half_extents, dtype=float)
        assert min(x, y, z) > 0
        verts = [
            [x, y, z], [x, y, -z], [x, -y, -z], [x, -y, z],
            [-x, y, z], [-x, y, -z], [-x, -y, -z], [-x, -y, z],
        ]
        return cls(verts, np.vstack((np.eye(3), -np.eye(3))))

    @classmethod
    def wedge(cls, half_extents):
        """Half a box cut along the diagonal; the slope faces +x."""
        hx, hy, hz = np.asarray(half_extents, dtype=float)
        assert min(hx, hy, hz) > 0
        verts = np.array(
            [[-hx, -hy, -hz], [hx, -hy, -hz], [-hx, -hy, hz],
             [-hx, hy, -hz], [hx, hy, -hz], [-hx, hy, hz]]
        )
        slope = np.cross(verts[4] - verts[1], verts[2] - verts[1])
        slope /= np.linalg.norm(slope)
        return cls(verts, np.vstack((-np.eye(3), [0, 1, 0], slope)))

    def transform(self, translation=None, rotation=None):
        t = np.zeros(3) if translation is None else np.asarray(translation, dtype=float)
        R = np.eye(3) if rotation is None else np.asarray(rotation, dtype=float)
        return ConvexPolyhedron(
            t + self.vertices @ R.T, self.normals @ R.T, R @ self.position + t, R @ self.rotation
        )

    def limits_along_axis(self, axis):
        proj = self.vertices @ (axis / np.linalg.norm(axis))
        return np.array([proj.min(), proj.max()])

    def max_vertex_along_axis(self, axis):
        proj = self.vertices @ (axis / np.linalg.norm(axis))
        return self.vertices[int(np.argmax(proj))]

    def height(self):
        lo, hi = self.limits_along_axis(np.array([0.0, 0.0, 1.0]))
        return hi - lo

    def polygon_in_plane(self, point, normal, span, tol=TOL):
        """CCW-wound 2-D polygon of the vertices lying in the plane."""
        on_plane = np.abs((self.vertices - point) @ normal) < tol
        return wind_ccw((self.vertices[on_plane] - point) @ span.T)[0]

    def distance_from_centroid_to_boundary(self, axis, offset=None):
        """Largest d with position+offset+d*axis still inside the hull.

        Solved as the same LP over convex-combination weights the reference
        uses (polyhedron.py:203-239).
        """
        axis = np.asarray(axis, dtype=float) / np.linalg.norm(axis)
        off = np.zeros(3) if offset is None else np.asarray(offset, dtype=float)
        n = self.vertices.shape[0]
        cost = np.zeros(n + 1)
        cost[0] = -1.0
        A = np.zeros((4, n + 1))
        A[:3, 0] = axis
        A[:3, 1:] = -self.vertices.T
        A[3, 1:] = 1.0
        b = np.ones(4)
        b[:3] = -(self.position + off)
        res = linprog(cost, A_eq=A, b_eq=b)
        d = res.x[0]
        assert d >= -TOL, "distance to boundary is negative"
        return d


def wind_ccw(V):
    """Sort 2-D points by angle about their mean; returns (V_sorted, index)."""
    V = np.asarray(V, dtype=float)
    c = V.mean(axis=0)
    idx = np.argsort(np.arctan2(V[:, 1] - c[1], V[:, 0] - c[0]))
    return V[idx], idx


def _clip_edge(v1, v2, point, normal, tol):
    """Part of segment v1→v2 on the positive side of the half-plane."""
    d1 = normal @ (v1 - point)
    d2 = normal @ (v2 - point)
    if d1 >= -tol and d2 >= -tol:
        return [v1, v2]
    if d1 <= tol and d2 <= tol:
        return []
    if abs(d1) < tol:
        hit = v1
    elif abs(d2) < tol:
        hit = v2
    else:
        t = normal @ (point - v1) / (normal @ (v2 - v1))
        hit = v1 + t * (v2 - v1)
    return [v1, hit] if d1 > 0 else [hit, v2]


def clip_polygon_with_half_space(V, point, normal, tol=TOL):
    normal = normal / np.linalg.norm(normal)
    pieces = []
    for i in range(V.shape[0]):
        pieces.extend(_clip_edge(V[i], V[(i + 1) % V.shape[0]], point, normal, tol))
    if not pieces:
        return None
    unique = []
    for p in pieces:
        if not any(np.linalg.norm(p - q) < tol for q in unique):
            unique.append(p)
    return np.array(unique)


def clip_polygon_with_polygon(V1, V2, tol=TOL):
    """Intersection of two CCW convex polygons (Sutherland–Hodgman)."""
    V = V1
    n = V2.shape[0]
    for i in range(n):
        edge = V2[(i + 1) % n] - V2[i]
        length = np.linalg.norm(edge)
        if length < tol:
            raise ValueError("clipping polygon has repeated vertices")
        inward = np.array([-edge[1], edge[0]]) / length
        V = clip_polygon_with_half_space(V, V2[i], inward, tol)
        if V is None:
            return None
    return V


def axis_aligned_contact(poly1, poly2, tol=TOL):
    """Contact manifold between two touching convex polyhedra.

    Returns (points (m,3), normal) with the normal pointing into `poly1`, or
    (None, None) if the shapes are separated or penetrating
    (polyhedron.py:446-514).  Candidate axes: face normals of both shapes then
    normalised cross products of face-normal pairs, in that order; the last
    axis on which the projections just touch defines the contact plane.
    """
    axes = [n for n in poly1.normals] + [n for n in poly2.normals]
    for n1 in poly1.normals:
        for n2 in poly2.normals:
            c = np.cross(n1, n2)
            mag = np.linalg.norm(c)
            if mag > tol:
                axes.append(c / mag)

    chosen = None
    for axis in axes:
        l1 = poly1.limits_along_axis(axis)
        l2 = poly2.limits_along_axis(axis)
        upper = min(l1[1], l2[1])
        lower = max(l1[0], l2[0])
        if abs(upper - lower) < tol:
            if l1[0] < l2[0]:
                chosen = (axis, poly1.max_vertex_along_axis(axis), -1.0)
            else:
                chosen = (axis, poly2.max_vertex_along_axis(axis), 1.0)
        elif upper < lower:
            return None, None
    if chosen is None:
        return None, None

    plane_normal, point, sign = chosen
    span = plane_span(plane_normal)
    V1 = poly1.polygon_in_plane(point, plane_normal, span, tol)
    V2 = poly2.polygon_in_plane(point, plane_normal, span, tol)
    overlap = clip_polygon_with_polygon(V1, V2, tol)
    if overlap is None:
        return None, None
    return point + overlap @ span, sign * plane_normal
