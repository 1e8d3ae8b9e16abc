"""Oracle: URDF parsing and batched forward kinematics in numpy (fp32 arithmetic, float64 constants cast to fp32).

Follows lib/utils/urdfpytorch/utils.py:22-51,142-167 (rpy, origin), urdf.py:2345-2398,2429-2464 (joint motion),
urdf.py:3064-3167 (per-link FK: fk[child] = fk[parent] @ (origin @ motion)), urdf.py:3813-3831,3920-3963 (column i of
the configuration <-> i-th actuated joint by stable ascending depth), lib/utils/urdf_robot.py:62-135,193-223
(keypoint links/offsets, get_keypoints, get_keypoints_root with the re-rooting inverse) and
lib/utils/geometries.py:100-115 (rot6d -> rotation matrix). Deliberately independent of the product's URDF compiler
(holistic-robot-pose-estimation-study_b200/urdf.py): it walks every link the way the reference does.
"""
import xml.etree.ElementTree as ET

import numpy as np

F32 = np.float32


def rpy_to_matrix(rpy):
    c3, c2, c1 = np.cos(np.asarray(rpy, np.float64))
    s3, s2, s1 = np.sin(np.asarray(rpy, np.float64))
    return np.array([[c1 * c2, c1 * s2 * s3 - c3 * s1, s1 * s3 + c1 * c3 * s2],
                     [c2 * s1, c1 * c3 + s1 * s2 * s3, c3 * s1 * s2 - c1 * s3],
                     [-s2, c2 * s3, c2 * c3]], np.float64)


class OracleRobot:
    def __init__(self, urdf_text):
        root = ET.fromstring(urdf_text)
        self.links = [e.attrib["name"] for e in root.findall("link")]
        self.joints = []
        for e in root.findall("joint"):
            M = np.eye(4)
            o = e.find("origin")
            if o is not None:
                if "xyz" in o.attrib:
                    M[:3, 3] = np.array(o.attrib["xyz"].split(), np.float64)
                if "rpy" in o.attrib:
                    M[:3, :3] = rpy_to_matrix(np.array(o.attrib["rpy"].split(), np.float64))
            a = e.find("axis")
            m = e.find("mimic")
            self.joints.append(dict(
                name=e.attrib["name"], type=e.attrib["type"], parent=e.find("parent").attrib["link"],
                child=e.find("child").attrib["link"], origin=M,
                axis=None if a is None else np.array(a.attrib["xyz"].split(), np.float64),
                mimic=None if m is None else (m.attrib["joint"], float(m.attrib.get("multiplier", 1.0)),
                                              float(m.attrib.get("offset", 0.0)))))
        self.by_child = {j["child"]: j for j in self.joints}
        self.by_name = {j["name"]: j for j in self.joints}
        self.base = [l for l in self.links if l not in self.by_child][0]
        act = [j for j in self.joints if j["type"] != "fixed" and j["mimic"] is None]
        depth = [len(self.path_to_base(j["child"])) for j in act]
        self.actuated = [act[i] for i in np.argsort(depth, kind="stable")]
        self.col = {j["name"]: i for i, j in enumerate(self.actuated)}
        # base-outward link order
        self.order = sorted(self.links, key=lambda l: len(self.path_to_base(l)))

    def path_to_base(self, link):
        p = [link]
        while link != self.base:
            link = self.by_child[link]["parent"]
            p.append(link)
        return p

    def child_pose(self, j, q):
        """[N,4,4] fp32 pose of the child frame in the parent frame."""
        n = q.shape[0]
        origin = j["origin"].astype(F32)
        if j["type"] == "fixed":
            return np.tile(origin, (n, 1, 1))
        if j["mimic"] is not None:
            src, mul, off = j["mimic"]
            cfg = (mul * q[:, self.col[src]] + off).astype(F32)
        else:
            cfg = q[:, self.col[j["name"]]]
        M = np.tile(np.eye(4, dtype=F32), (n, 1, 1))
        if j["type"] in ("revolute", "continuous"):
            axis = j["axis"] / np.linalg.norm(j["axis"])
            s, c = np.sin(cfg), np.cos(cfg)
            M[:, 0, 0] = c
            M[:, 1, 1] = c
            M[:, 2, 2] = c
            M[:, :3, :3] += np.outer(axis, axis).astype(F32)[None] * (F32(1.0) - c)[:, None, None]
            K = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
            M[:, :3, :3] += K.astype(F32)[None] * s[:, None, None]
        else:  # prismatic
            M[:, :3, 3] = j["axis"].astype(F32)[None] * cfg[:, None]
        return np.matmul(origin[None], M)

    def link_fk(self, q):
        """dict link -> [N,4,4] fp32 pose in the base frame."""
        q = np.asarray(q, F32)
        fk = {self.base: np.tile(np.eye(4, dtype=F32), (q.shape[0], 1, 1))}
        for l in self.order:
            if l == self.base:
                continue
            j = self.by_child[l]
            fk[l] = np.matmul(fk[j["parent"]], self.child_pose(j, q))
        return fk


def rot6d_to_rotmat(r):
    """geometries.py:100-115: rows (x, y, z); x = a1/|a1|, z = (x x a2)/|.|, y = z x x. No epsilon."""
    r = np.asarray(r, F32)
    x = r[:, 0:3] / np.linalg.norm(r[:, 0:3], axis=1, keepdims=True)
    z = np.cross(x, r[:, 3:6])
    z = z / np.linalg.norm(z, axis=1, keepdims=True)
    y = np.cross(z, x)
    return np.stack((x, y, z), 1).astype(F32)


def keypoints(robot, kp_frames, q, rot6d, trans, root=0):
    """urdf_robot.py:95-118 (root == 0) / 193-223 (root > 0): [N, nkpt, 3] camera-frame keypoints."""
    n = q.shape[0]
    fk = robot.link_fk(q)
    T = np.stack([fk[l] for l, _ in kp_frames], 1)                           # [N, K, 4, 4]
    off = np.stack([np.asarray(o, np.float64) for _, o in kp_frames]).astype(F32)
    b2c = np.zeros((n, 4, 4), F32)
    b2c[:, :3, :3] = rot6d_to_rotmat(rot6d)
    b2c[:, :3, 3] = trans
    b2c[:, 3, 3] = 1.0
    if root != 0:
        T = np.matmul(np.linalg.inv(T[:, root:root + 1]).astype(F32), T)
    T = np.matmul(b2c[:, None], T)
    return (np.matmul(T[:, :, :3, :3], off[None, :, :, None])[..., 0] + T[:, :, :3, 3]).astype(F32)
