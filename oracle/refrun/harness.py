"""Run the UNMODIFIED reference (/root/reference) on CPU in this container.

TEST INFRASTRUCTURE ONLY. Used by the golden-vector generator scripts (oracle/refrun/make_golden*.py) to pin the
oracle port (oracle/*.py) and, through the committed fixtures in tests/golden/, the CUDA path. `/root/reference` does not
exist on the GPU box, so nothing here may be imported by `-m gpu` tests or `smoke()`; bench.py's reference arm / cpu_baseline
leg (the one place allowed to execute oracle/ besides the tests) runs it from the travelling copy baseline/_ref/ (see REF).

What it does (recipe from SURVEY.md §8c, nothing in the reference tree is modified or copied):
  * builds a scratch working directory (`data/`, `lib -> /root/reference/lib`, kinematics-only URDFs at the paths
    lib/config.py:33-36 expects) and chdir()s into it (lib/config.py:21 asserts ./data, HRnet.py:614 opens a cwd-relative
    YAML);
  * puts stdlib shims for absent packages first on sys.path (easydict, lxml.etree, trimesh, pyrender, pytorch3d,
    roboticstoolbox) -- none of them touch arithmetic;
  * patches `torch.Tensor.cuda` to the identity (the reference hard-codes .cuda(), integral.py:73, transforms.py:54...),
    the two `init_weights` (network/file access, Resnet.py:70, HRnet.py:572) to no-ops, and `URDF._sort_joints` to a
    STABLE argsort (urdf.py:3830 relies on numpy<=1.22 tie behaviour; numpy 2.x permutes Baxter's equal-depth joints).
"""
import argparse
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
# the reference tree: the read-only original in the build container, else the git-ignored copy of its lib/ + configs/ that
# __graft_entry__.build() leaves under baseline/_ref/ (it travels to the GPU box with the snapshot; bench.py --impl reference
# and the cpu_baseline leg time the reference's own forward from it). Never imported by the product package.
REF = "/root/reference" if os.path.isdir("/root/reference/lib") else os.path.join(REPO, "baseline", "_ref")
URDF_DIR = os.path.join(REPO, "holistic-robot-pose-estimation-study_b200", "data", "urdf")

_state = {}


def available():
    return os.path.isdir(os.path.join(REF, "lib")) and os.path.isdir(os.path.join(REF, "configs"))


def setup():
    """Idempotent. Returns a namespace with the reference modules."""
    if _state:
        return _state["ns"]
    if not available():
        raise RuntimeError("reference tree not present (neither /root/reference nor baseline/_ref)")
    sys.dont_write_bytecode = True
    work = tempfile.mkdtemp(prefix="hrp_refrun_")
    os.makedirs(os.path.join(work, "data", "deps", "panda-description", "patched_urdf"))
    os.makedirs(os.path.join(work, "data", "deps", "kuka-description", "iiwa_description", "urdf"))
    os.makedirs(os.path.join(work, "data", "deps", "baxter-description", "baxter_description", "urdf"))
    os.symlink(os.path.join(REF, "lib"), os.path.join(work, "lib"))
    os.symlink(os.path.join(REF, "configs"), os.path.join(work, "configs"))
    d = os.path.join(work, "data", "deps")
    shutil.copy(os.path.join(URDF_DIR, "panda.urdf"), os.path.join(d, "panda-description", "panda.urdf"))
    shutil.copy(os.path.join(URDF_DIR, "panda.urdf"), os.path.join(d, "panda-description", "patched_urdf", "panda.urdf"))
    shutil.copy(os.path.join(URDF_DIR, "iiwa7.urdf"),
                os.path.join(d, "kuka-description", "iiwa_description", "urdf", "iiwa7.urdf"))
    baxter = os.path.join(d, "baxter-description", "baxter_description", "urdf", "baxter.urdf")
    shutil.copy(os.path.join(URDF_DIR, "baxter.urdf"), baxter)
    os.chdir(work)
    sys.path[:0] = [os.path.join(HERE, "shims"), os.path.join(REF, "lib"), REF]

    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU oracle only

    import config as ref_config  # lib/config.py
    ref_config.BAXTER_DESCRIPTION_PATH = baxter
    import lib.config as ref_config2  # same file under its second import name (core/config.py:5)
    ref_config2.BAXTER_DESCRIPTION_PATH = baxter

    from utils.urdfpytorch import urdf as ref_urdf

    def _stable_sort_joints(self, joints):
        lens = [len(self._paths_to_base[self._link_map[j.child]]) for j in joints]
        order = np.argsort(lens, kind="stable")
        return np.array(joints)[order].tolist()

    ref_urdf.URDF._sort_joints = _stable_sort_joints

    from models.backbones import HRnet as ref_hrnet, Resnet as ref_resnet
    ref_hrnet.PoseHighResolutionNet.init_weights = lambda self, *a, **k: None
    ref_resnet.ResNet.init_weights = lambda self, *a, **k: None

    from models import full_net as ref_full_net
    from utils import transforms as ref_transforms, integral as ref_integral, geometries as ref_geom
    from utils import urdf_robot as ref_urdf_robot
    from dataset import const as ref_const
    from core import config as ref_core_config

    ns = argparse.Namespace(work=work, full_net=ref_full_net, transforms=ref_transforms, integral=ref_integral,
                            geometries=ref_geom, urdf_robot=ref_urdf_robot, const=ref_const,
                            core_config=ref_core_config, hrnet=ref_hrnet, resnet=ref_resnet)
    _state["ns"] = ns
    return ns


def make_cfg(robot, backbone_name=None, **overrides):
    """args object the reference ctor reads (full_net.py:58-75), from the shipped YAML (configs/<robot>/full.yaml);
    `overrides` set constructor switches the shipped files leave at their defaults (direct_reg_rot, add_fc, multi_kp, ...)."""
    ns = setup()
    cfg = ns.core_config.make_cfg(argparse.Namespace(config=f"configs/{robot}/full.yaml", resume=False))
    cfg.pretrained_rootnet = None
    if backbone_name is not None:
        cfg.backbone_name = backbone_name
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def build_model(robot, backbone_name=None, **overrides):
    """Construct RootNetwithRegInt directly (the factory's init_weights call raises for hrnet32)."""
    ns = setup()
    cfg = make_cfg(robot, backbone_name, **overrides)
    init = {"robot_type": robot, "pose_params": ns.const.INITIAL_JOINT_ANGLE,
            "cam_params": np.eye(4, dtype=float), "init_pose_from_mean": True}
    model = ns.full_net.RootNetwithRegInt(init, cfg)
    model.eval()
    for r in (model.robot.robot,):
        names = [j.name for j in r.actuated_joints]
        assert names == ns.const.JOINT_NAMES[robot], (names, ns.const.JOINT_NAMES[robot])
    return model, cfg


def forward(model, x_reg, x_root, k_value, K):
    """The boundary call (function.py:133-141): 8-tuple + caller-side projections of both 3-D keypoint sets."""
    ns = setup()
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk"]
    with torch.no_grad():
        out = model(x_reg, x_root, k_value, K)
        if len(out) == 9:                     # multi_kp: every regressed depth rides along after the root depth (full_net.py:462-464)
            names = names[:5] + ["depths"] + names[5:]
        res = {n: o for n, o in zip(names, out)}
        kp2d_int = ns.transforms.point_projection_from_3d_tensor(K, res["kp3d_int"])
        kp2d_fk = ns.transforms.point_projection_from_3d_tensor(K, res["kp3d_fk"])
    res["kp2d_int"] = kp2d_int
    res["kp2d_fk"] = kp2d_fk
    return res
