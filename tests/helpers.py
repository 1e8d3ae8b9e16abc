"""Shared helpers for the parity tests (inputs are re-derived from seeds recorded in the golden fixtures)."""
import os

import numpy as np
import torch

from hrp_b200 import consts, synth
from oracle import model as omodel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULLNET_CASES = [("panda", "resnet50"), ("kuka", "resnet50"), ("baxter", "resnet50"), ("panda", "hrnet32"),
                 ("baxter", "hrnet32"), ("kuka", "hrnet32")]
UNDAMPED_CASES = [("panda", "resnet50"), ("panda", "hrnet32")]
# north_star parity gates (fp32 / TF32 parity mode)
TOL_PX, TOL_RAD, TOL_DEPTH_M = 0.5, 1e-3, 1e-3

_oracles = {}


def oracle_for(robot, backbone, seed=1234, recipe="damped"):
    key = (robot, backbone, seed, recipe)
    if key not in _oracles:
        sd = synth.make_state_dict(robot, backbone, seed, recipe=recipe)
        _oracles[key] = (omodel.OracleModel(robot, sd, open(consts.urdf_path(robot)).read(), backbone), sd)
    return _oracles[key]


# constructor variants outside the shipped configuration (tests/golden/variant_<name>.npz, made by the reference built with
# these switches: oracle/refrun/make_golden.py variants): name -> (config keys as the reference spells them, ctor dict of the port)
VARIANT_CASES = {
    "direct_addfc_multikp": (dict(direct_reg_rot=True, add_fc=True, multi_kp=True, kps_need_depth=[0, 3, 6]),
                             dict(direct_reg_rot=True, rot_iterative_matmul=False, add_fc=True, depth_num=3, depth_root=1)),
    "rotmatmul": (dict(rot_iterative_matmul=True),
                  dict(direct_reg_rot=False, rot_iterative_matmul=True, add_fc=False, depth_num=1, depth_root=0)),
    "jointmap": (dict(reg_joint_map=True, joint_conv_dim=[128, 64, 32]),
                 dict(reg_joint_map=True, joint_conv_dim=(128, 64, 32), joint_bounds=consts.ROBOTS["panda"]["bounds"])),
}
JOINTMAP_GAIN = 6.0        # oracle/refrun/make_golden.py: contrast of the joint maps in the fixture


def variant_state_dict(name, seed=1234):
    ctor = VARIANT_CASES[name][1]
    sd = synth.make_state_dict("panda", "resnet50", seed, ctor={k: v for k, v in ctor.items() if k not in ("depth_root", "joint_bounds")})
    if ctor.get("reg_joint_map"):
        sd["joint_final_layer.weight"] = sd["joint_final_layer.weight"] * np.float32(JOINTMAP_GAIN)
    return sd


def checkpoint_case():
    """The weights of tests/golden/fullnet_panda_resnet50_ckpt.npz as the two reference-format checkpoints it was made
    from (oracle/refrun/make_golden.py checkpoint()): (main checkpoint dict with DataParallel-prefixed keys and no DepthNet
    tensors, DepthNet pre-training checkpoint dict in lib/models/depth_net.py naming, merged plain state dict)."""
    import collections
    sd = synth.make_state_dict("panda", "resnet50", 1234, recipe="undamped")
    pre = synth.make_pretrained_rootnet_state(sd)
    main = collections.OrderedDict(("module." + k, torch.from_numpy(np.asarray(v))) for k, v in sd.items()
                                   if not k.startswith(("rootnet_backbone.", "depth_layer.")))
    merged = dict(sd)
    for k, v in pre.items():
        nk = k.replace("backbone", "rootnet_backbone") if k.startswith("backbone") else k
        if nk in merged:
            merged[nk] = v
    ck_main = {"epoch": 7, "auc_add": 0.5, "model_state_dict": main, "optimizer_state_dict": {}, "lr_scheduler_last_epoch": -1}
    ck_pre = {"epoch": 3, "model_state_dict": collections.OrderedDict((k, torch.from_numpy(v)) for k, v in pre.items())}
    return ck_main, ck_pre, merged


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def inputs(B, seed):
    img, K, kv = synth.make_inputs(B, seed)
    return torch.from_numpy(img), torch.from_numpy(K), torch.from_numpy(kv)


def maxdiff(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
