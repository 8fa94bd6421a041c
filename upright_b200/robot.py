"""Kinematic-chain fixture for the "Thing" (Ridgeback + UR10) and the
fixed-base UR10, plus a small numpy forward kinematics used on the host for
reference poses (the role `upright_control/src/upright_control/robot.py:92-140`
gives to pinocchio-python).

What the reference fixes (and is followed here):
  * root joint = prismatic x, prismatic y, revolute z, world-frame velocities
    (upright_control/include/upright_control/util.h:27-32, robot.py:12-16);
  * fixed base = root locked at `base_pose` (util.h:34-48);
  * joint order of q (upright_cmd/config/robots/thing.yaml:15-24);
  * tray frame `gripped_object` relative to link `gripper`
    (upright_assets/thing/xacro/end_effectors/gripped_object.urdf.xacro:6-10 with
    upright_cmd/config/robots/calibration/tray_transforms_sim_2023-01-09_13-07-23.yaml);
  * collision spheres (upright_assets/thing/xacro/collision_links.urdf.xacro:31-184).

What the reference does NOT contain (un-vendored `mobile_manipulation_central`
URDF `thing_no_wheels.urdf.xacro`, thing.yaml:32,53) and is therefore a
documented fixture of this repository:
  * UR10 link offsets: the public Universal Robots UR10 description values;
  * the arm mount pose on the Ridgeback and the `gripper` frame on the flange
    (`ARM_MOUNT_*`, `GRIPPER_YAW` below; the latter is chosen so that the tray
    is exactly level at the shipped home configuration thing.yaml:48).
Parity is defined against the in-repo oracle on this same fixture.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .geometry import rotz, rpy_to_rot, skew3

REVOLUTE, PRISMATIC = 0, 1

# --- public UR10 DH parameters (Universal Robots / ur_description default kinematics) ---
UR10_D1 = 0.1273
UR10_A2 = -0.612
UR10_A3 = -0.5723
UR10_D4 = 0.163941
UR10_D5 = 0.1157
UR10_D6 = 0.0922

# --- fixture assumptions (not in the reference tree) ---
ARM_MOUNT_XYZ = (0.27, 0.01, 0.653)
ARM_MOUNT_RPY = (0.0, 0.0, -0.5 * np.pi)
# tool0 -> gripper: rotation about the flange axis.  0.417*pi (wrist_3 at
# home) + GRIPPER_YAW = pi/2 makes the gripped tray level at home.
GRIPPER_YAW = (0.5 - 0.417) * np.pi

# gripper -> gripped_object (sim calibration file, nominal values)
TRAY_RPY = (0.0, -1.57079633, 3.14159265)
TRAY_XYZ = (0.036712437868118286, -0.0004053786105941981, 0.308562308549881)


@dataclass
class Joint:
    name: str
    type: int
    R: np.ndarray  # parent link -> joint frame
    p: np.ndarray
    axis: np.ndarray


@dataclass
class Sphere:
    name: str
    link: int  # 0..nq-1 link after that joint, nq = tool frame, -1 = world
    offset: np.ndarray
    radius: float
    shape: int = 0  # 0 sphere; 1 half-space (offset = unit normal, radius = plane offset): the reference's `ground`


@dataclass
class KinematicChain:
    joints: list
    tool_R: np.ndarray
    tool_p: np.ndarray
    link_names: list = field(default_factory=list)
    spheres: list = field(default_factory=list)

    @property
    def nq(self):
        return len(self.joints)

    # -- numpy FK (positions/orientations only) --
    def link_frames(self, q):
        R, p = np.eye(3), np.zeros(3)
        frames = []
        for j, qi in zip(self.joints, q):
            p = p + R @ j.p
            R = R @ j.R
            if j.type == REVOLUTE:
                R = R @ _axis_angle(j.axis, qi)
            else:
                p = p + R @ j.axis * qi
            frames.append((R.copy(), p.copy()))
        return frames

    def tool_pose(self, q):
        R, p = self.link_frames(q)[-1]
        return p + R @ self.tool_p, R @ self.tool_R

    def sphere_centers(self, q):
        frames = self.link_frames(q)
        Rt, pt = None, None
        out = []
        for s in self.spheres:
            if s.link < 0:
                out.append(np.array(s.offset, dtype=float))
            elif s.link == self.nq:
                if Rt is None:
                    pt, Rt = self.tool_pose(q)
                out.append(pt + Rt @ s.offset)
            else:
                R, p = frames[s.link]
                out.append(p + R @ s.offset)
        return np.array(out)


def _axis_angle(axis, angle):
    K = skew3(axis)
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def _tool_transform(tray_rpy=TRAY_RPY, tray_xyz=TRAY_XYZ):
    # wrist_3_link (== tool0 in the DH-style convention: z out of the flange)
    # -> gripper -> gripped_object
    R = np.eye(3)
    p = np.zeros(3)
    R = R @ rotz(GRIPPER_YAW)
    p = p + R @ np.asarray(tray_xyz, dtype=float)
    R = R @ rpy_to_rot(tray_rpy)
    return R, p


def _arm_joints(first_R, first_p):
    """Six UR10 revolute joints in the DH-style frame convention of the current
    ur_description (every joint turns about its local z; links run along -x),
    which is the convention the reference's collision-sphere offsets assume
    (forearm spheres at x = -0.2 / -0.4, collision_links.urdf.xacro:160-184).
    `first_R/first_p` carry base/mount -> base_link_inertia (incl. its pi yaw)
    -> shoulder joint origin."""
    ez = np.array([0.0, 0.0, 1.0])
    I3 = np.eye(3)
    rx90 = rpy_to_rot((0.5 * np.pi, 0.0, 0.0))
    return [
        Joint("ur10_arm_shoulder_pan_joint", REVOLUTE, first_R, first_p, ez),
        Joint("ur10_arm_shoulder_lift_joint", REVOLUTE, rx90, np.zeros(3), ez),
        Joint("ur10_arm_elbow_joint", REVOLUTE, I3, np.array([UR10_A2, 0.0, 0.0]), ez),
        Joint("ur10_arm_wrist_1_joint", REVOLUTE, I3, np.array([UR10_A3, 0.0, UR10_D4]), ez),
        Joint("ur10_arm_wrist_2_joint", REVOLUTE, rx90, np.array([0.0, -UR10_D5, 0.0]), ez),
        Joint("ur10_arm_wrist_3_joint", REVOLUTE, rpy_to_rot((0.5 * np.pi, np.pi, np.pi)),
              np.array([0.0, UR10_D6, 0.0]), ez),
    ]


_ARM_LINKS = [
    "ur10_arm_shoulder_link", "ur10_arm_upper_arm_link", "ur10_arm_forearm_link",
    "ur10_arm_wrist_1_link", "ur10_arm_wrist_2_link", "ur10_arm_wrist_3_link",
]


def _robot_spheres(link_index, base_link, base_offset=None):
    """Collision spheres of collision_links.urdf.xacro:31-184."""
    base_off = np.zeros(3) if base_offset is None else base_offset
    return [
        Sphere("balanced_object_collision_link", link_index["gripped_object"], np.array([0, 0, 0.07]), 0.25),
        Sphere("shoulder_collision_link", link_index["ur10_arm_upper_arm_link"], np.zeros(3), 0.15),
        Sphere("wrist1_collision_link", link_index["ur10_arm_wrist_1_link"], np.array([0, 0, -0.05]), 0.15),
        Sphere("wrist3_collision_link", link_index["ur10_arm_wrist_3_link"], np.zeros(3), 0.15),
        Sphere("base_collision_link", base_link, base_off, 0.5),
        Sphere("forearm_collision_sphere_link1", link_index["ur10_arm_forearm_link"], np.array([-0.2, 0, 0.06]), 0.15),
        Sphere("forearm_collision_sphere_link2", link_index["ur10_arm_forearm_link"], np.array([-0.4, 0, 0.06]), 0.15),
    ]


def build_chain(base_type="omnidirectional", base_pose=(0.0, 0.0, 0.0),
                tray_rpy=TRAY_RPY, tray_xyz=TRAY_XYZ) -> KinematicChain:
    """Thing (`omnidirectional`, nq=9) or UR10 on a locked base (`fixed`, nq=6)."""
    ex, ey, ez = np.eye(3)
    # base_link -> ur10_arm_base_link -> base_link_inertia (pi about z) -> +d1
    mount_R = rpy_to_rot(ARM_MOUNT_RPY) @ rotz(np.pi)
    mount_p = np.asarray(ARM_MOUNT_XYZ, dtype=float) + np.array([0.0, 0.0, UR10_D1])
    tool_R, tool_p = _tool_transform(tray_rpy, tray_xyz)
    if base_type == "omnidirectional":
        root = [
            Joint("x_to_world_joint", PRISMATIC, np.eye(3), np.zeros(3), ex),
            Joint("y_to_x_joint", PRISMATIC, np.eye(3), np.zeros(3), ey),
            Joint("base_to_y_joint", REVOLUTE, np.eye(3), np.zeros(3), ez),
        ]
        joints = root + _arm_joints(mount_R, mount_p)
        links = ["x_link", "y_link", "base_link"] + _ARM_LINKS
        index = {n: i for i, n in enumerate(links)}
        index["gripped_object"] = len(joints)
        spheres = _robot_spheres(index, index["base_link"])
    elif base_type == "fixed":
        bx, by, bth = base_pose
        Rb = rotz(bth)
        pb = np.array([bx, by, 0.0])
        joints = _arm_joints(Rb @ mount_R, pb + Rb @ mount_p)
        links = list(_ARM_LINKS)
        index = {n: i for i, n in enumerate(links)}
        index["gripped_object"] = len(joints)
        spheres = _robot_spheres(index, -1, pb)
    else:
        raise ValueError(f"unsupported base type {base_type!r}")
    return KinematicChain(joints, tool_R, tool_p, links + ["gripped_object"], spheres)
