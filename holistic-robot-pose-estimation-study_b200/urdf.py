"""URDF text -> flat kinematic program for the batched FK kernel (csrc/fk_project.cu).

Host-side, runs once per robot at handle creation. Semantics follow the reference's URDF loader and batched FK:
origin = xyz + Z1-Y2-X3 rpy (lib/utils/urdfpytorch/utils.py:22-51,142-167), revolute = origin . Rodrigues(axis, q)
(urdf.py:2385-2386, 2429-2464), prismatic = origin . translate(axis * q) (2387-2392), mimic q' = multiplier*q + offset
(3143-3148), configuration column i <-> i-th actuated joint in STABLE ascending depth order (3813-3831, 3950-3953),
keypoint = link origin (+ constant offset in the link frame; Baxter: child-joint origins in the parent link,
lib/utils/urdf_robot.py:62-87).

The program is a compiled form, not a transliteration: only joints on a path to some keypoint link survive, runs of
fixed joints are folded (in float64) into the next movable joint's origin or into the keypoint offset, and each step
says whether it continues the previous step's transform, restarts from the base, or reloads a saved transform.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

from . import consts

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
PARENT_BASE, PARENT_PREV = -1, -2

MAX_STEPS = 32
MAX_KP = 32
MAX_SLOTS = 8


def _rpy_matrix(r, p, y):
    cr, cp, cy = math.cos(r), math.cos(p), math.cos(y)
    sr, sp, sy = math.sin(r), math.sin(p), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - cr * sy, sy * sr + cy * cr * sp],
                     [cp * sy, cy * cr + sy * sp * sr, cr * sy * sp - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


def _floats(s):
    return [float(v) for v in s.split()]


class Joint:
    __slots__ = ("name", "type", "parent", "child", "origin", "axis", "mimic", "index")


class Robot:
    """Parsed URDF: links, joints, parent map, depth-sorted actuated joints."""

    def __init__(self, text):
        root = ET.fromstring(text)
        self.name = root.attrib.get("name", "")
        self.links = [l.attrib["name"] for l in root.findall("link")]
        self.joints = []
        for j in root.findall("joint"):
            J = Joint()
            J.name = j.attrib["name"]
            J.type = j.attrib["type"]
            J.parent = j.find("parent").attrib["link"]
            J.child = j.find("child").attrib["link"]
            M = np.eye(4, dtype=np.float64)
            o = j.find("origin")
            if o is not None:
                if "xyz" in o.attrib:
                    M[:3, 3] = _floats(o.attrib["xyz"])
                if "rpy" in o.attrib:
                    M[:3, :3] = _rpy_matrix(*_floats(o.attrib["rpy"]))
            J.origin = M
            a = j.find("axis")
            J.axis = np.array(_floats(a.attrib["xyz"]) if a is not None else [1.0, 0.0, 0.0], dtype=np.float64)
            m = j.find("mimic")
            J.mimic = None
            if m is not None:
                J.mimic = (m.attrib["joint"], float(m.attrib.get("multiplier", 1.0)), float(m.attrib.get("offset", 0.0)))
            if J.type not in ("fixed", "revolute", "continuous", "prismatic"):
                raise ValueError("unsupported joint type %s (%s)" % (J.type, J.name))
            self.joints.append(J)
        self.joint_by_name = {j.name: j for j in self.joints}
        self.joint_of_child = {j.child: j for j in self.joints}
        children = set(self.joint_of_child)
        bases = [l for l in self.links if l not in children]
        if len(bases) != 1:
            raise ValueError("URDF must have exactly one base link, got %r" % bases)
        self.base = bases[0]
        actuated = [j for j in self.joints if j.type != "fixed" and j.mimic is None]
        depth = [self.depth(j.child) for j in actuated]
        order = sorted(range(len(actuated)), key=lambda i: depth[i])  # stable
        self.actuated = [actuated[i] for i in order]
        for i, j in enumerate(self.actuated):
            j.index = i

    def depth(self, link):
        d = 1
        while link != self.base:
            link = self.joint_of_child[link].parent
            d += 1
        return d

    def path_from_base(self, link):
        p = []
        while link != self.base:
            j = self.joint_of_child[link]
            p.append(j)
            link = j.parent
        return p[::-1]


class KinematicProgram:
    """Flat tables consumed by `hrp_fk_create` (include/hrp_b200.h)."""

    def __init__(self):
        self.dof = 0
        self.nkpt = 0
        self.root_kp = 0            # keypoint index whose LINK frame re-roots the chain (0 = no re-rooting)
        self.root_step = -1         # step whose transform is the root link frame (-1: base)
        self.root_fixed = np.eye(4)[:3].astype(np.float32)  # constant transform from root_step frame to the root link
        self.step_type = []
        self.step_parent = []
        self.step_save = []
        self.step_q = []
        self.step_mul = []
        self.step_off = []
        self.step_origin = []       # 12 floats each
        self.step_axis = []         # 3 floats each
        self.kp_step = []           # sorted ascending; -1 = base frame
        self.kp_index = []          # output keypoint index
        self.kp_offset = []         # 3 floats each, in the step frame
        self.n_slots = 0

    def arrays(self):
        f32, i32 = np.float32, np.int32
        return dict(
            step_type=np.asarray(self.step_type, i32), step_parent=np.asarray(self.step_parent, i32),
            step_save=np.asarray(self.step_save, i32), step_q=np.asarray(self.step_q, i32),
            step_mul=np.asarray(self.step_mul, f32), step_off=np.asarray(self.step_off, f32),
            step_origin=np.asarray(self.step_origin, f32).reshape(-1, 12),
            step_axis=np.asarray(self.step_axis, f32).reshape(-1, 3),
            kp_step=np.asarray(self.kp_step, i32), kp_index=np.asarray(self.kp_index, i32),
            kp_offset=np.asarray(self.kp_offset, f32).reshape(-1, 3),
            root_fixed=np.asarray(self.root_fixed, f32).reshape(12),
        )


def keypoint_frames(robot_type, R):
    """(link name, offset xyz) per keypoint -- urdf_robot.py:62-87."""
    spec = consts.ROBOTS[robot_type]
    if "kp_joints" in spec:
        out = []
        for jn in spec["kp_joints"]:
            j = R.joint_by_name[jn]
            out.append((j.parent, j.origin[:3, 3].copy()))
        return out
    return [(l, np.zeros(3)) for l in spec["links"]]


def compile_program(R, kp_frames, root_kp, expected_joints=None):
    """Compile parsed URDF `R` + keypoint frames into a KinematicProgram."""
    if expected_joints is not None:
        names = [j.name for j in R.actuated]
        if names != list(expected_joints):
            raise ValueError("actuated joint order %r != expected %r" % (names, expected_joints))
    P = KinematicProgram()
    P.dof = len(R.actuated)
    P.nkpt = len(kp_frames)
    P.root_kp = root_kp
    if P.nkpt > MAX_KP:
        raise ValueError("too many keypoints")

    # movable joints needed by some keypoint, in a DFS order that keeps chains contiguous
    needed = []
    for link, _ in kp_frames:
        for j in R.path_from_base(link):
            if j.type != "fixed" and j not in needed:
                needed.append(j)

    def movable_ancestor(link):
        """(last movable joint on the path or None, fixed transform from that joint's child frame to `link`)."""
        F = np.eye(4)
        last = None
        for j in R.path_from_base(link):
            if j.type == "fixed":
                F = F @ j.origin
            else:
                last, F = j, np.eye(4)
        return last, F

    # order: depth-first over the tree of needed movable joints, so serial chains stay contiguous
    kids = {}
    for j in needed:
        pj, _ = movable_ancestor(j.parent)
        kids.setdefault(pj, []).append(j)
    order = []

    def place(j):
        order.append(j)
        for c in kids.get(j, []):
            place(c)

    for j in kids.get(None, []):
        place(j)
    if len(order) > MAX_STEPS:
        raise ValueError("too many kinematic steps")
    step_of = {j: i for i, j in enumerate(order)}

    # which steps must be saved (a later, non-adjacent step starts from them)
    slot_of = {}
    for i, j in enumerate(order):
        pj, _ = movable_ancestor(j.parent)
        if pj is not None and step_of[pj] != i - 1 and pj not in slot_of:
            slot_of[pj] = len(slot_of)
    if len(slot_of) > MAX_SLOTS:
        raise ValueError("kinematic tree needs too many saved frames")
    P.n_slots = len(slot_of)

    for i, j in enumerate(order):
        pj, F = movable_ancestor(j.parent)
        origin = F @ j.origin                      # fixed joints folded in float64
        if pj is None:
            parent = PARENT_BASE
        elif step_of[pj] == i - 1:
            parent = PARENT_PREV
        else:
            parent = slot_of[pj]
        src, mul, off = j, 1.0, 0.0
        if j.mimic is not None:
            src, mul, off = R.joint_by_name[j.mimic[0]], j.mimic[1], j.mimic[2]
        if j.type in ("revolute", "continuous"):
            typ = REVOLUTE
            axis = j.axis / np.linalg.norm(j.axis)  # urdf.py:2448
        else:
            typ = PRISMATIC
            axis = j.axis                           # used un-normalised, urdf.py:2391
        P.step_type.append(typ)
        P.step_parent.append(parent)
        P.step_save.append(slot_of.get(j, -1))
        P.step_q.append(src.index)
        P.step_mul.append(mul)
        P.step_off.append(off)
        P.step_origin.append(origin[:3, :].astype(np.float32).reshape(12))
        P.step_axis.append(axis.astype(np.float32))

    kps = []
    for k, (link, off) in enumerate(kp_frames):
        mj, F = movable_ancestor(link)
        o = (F @ np.append(off, 1.0))[:3]
        kps.append((-1 if mj is None else step_of[mj], k, o.astype(np.float32)))
    kps.sort(key=lambda t: (t[0], t[1]))
    for s, k, o in kps:
        P.kp_step.append(s)
        P.kp_index.append(k)
        P.kp_offset.append(o)

    if root_kp != 0:
        mj, F = movable_ancestor(kp_frames[root_kp][0])
        P.root_step = -1 if mj is None else step_of[mj]
        P.root_fixed = F[:3, :].astype(np.float32)
    return P


def load_robot(robot_type, text=None):
    """Parse the packaged URDF of `robot_type` and compile its keypoint program."""
    spec = consts.ROBOTS[robot_type]
    if text is None:
        with open(consts.urdf_path(robot_type)) as f:
            text = f.read()
    R = Robot(text)
    frames = keypoint_frames(robot_type, R)
    P = compile_program(R, frames, spec["ref_kp"], spec["joints"])
    assert P.dof == spec["dof"] and P.nkpt == spec["nkpt"]
    return R, P
