"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only; see harness.py).

Fixtures (inputs are re-derivable from the recorded seeds through hrp_b200.synth, so only outputs are stored):
  fullnet_<robot>_<backbone>.npz  reference forward 8-tuple + caller-side projections on synthetic weights/inputs,
                                  plus stage-boundary probes captured with forward hooks (xf, img_feat, a strided
                                  sample of the heatmap logits);
  softargmax_<path>_k<nkpt>_<mode>.npz   HeatmapIntegralPose (integral.py:102-208) on adversarial heatmaps;
  fk_<robot>.npz                  URDFRobot.get_keypoints[_root] (urdf_robot.py:95-118,193-223) +
                                  point_projection_from_3d_tensor (transforms.py:17-21) over the joint-bound sweep.
Usage: python -m oracle.refrun.make_golden
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
import hrp_b200  # noqa: E402,F401
from hrp_b200 import consts, synth  # noqa: E402
from oracle.refrun import harness  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
WEIGHT_SEED = 1234
FULLNET = [("panda", "resnet50", 2, 2024), ("kuka", "resnet50", 2, 2025), ("baxter", "resnet50", 2, 2026),
           ("panda", "hrnet32", 2, 2027), ("baxter", "hrnet32", 1, 2028)]
LOGIT_STRIDES = (37, 5, 7)
FK_N, FK_SEED = 512, 99
SA_CASES = [("resnet50", 7, "blobs", 2, 11), ("resnet50", 7, "extreme", 2, 12), ("resnet50", 17, "noise", 1, 13),
            ("hrnet32", 8, "blobs", 2, 14), ("hrnet32", 17, "extreme", 1, 15)]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def fullnet():
    for robot, bb, B, seed in FULLNET:
        sd = synth.make_state_dict(robot, bb, WEIGHT_SEED)
        model, _ = harness.build_model(robot, bb)
        model.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        probes = {}
        hooks = [model.rootnet_backbone.register_forward_hook(lambda m, i, o: probes.__setitem__("img_feat", o))]
        if bb == "resnet50":
            hooks.append(model.avgpool.register_forward_hook(lambda m, i, o: probes.__setitem__("xf", o.flatten(1))))
            hooks.append(model.final_layer.register_forward_hook(lambda m, i, o: probes.__setitem__("logits", o)))
        else:
            def grab(m, i, o):
                probes["logits"], probes["xf"] = o[0], o[1]
            hooks.append(model.reg_backbone.register_forward_hook(grab))
        img, K, kv = synth.make_inputs(B, seed)
        res = harness.forward(model, t(img), t(img), t(kv), t(K))
        for h in hooks:
            h.remove()
        s = LOGIT_STRIDES
        out = {k: v.numpy() for k, v in res.items()}
        out["probe_xf"] = probes["xf"].numpy()
        out["probe_img_feat"] = probes["img_feat"].numpy()
        lg = probes["logits"]
        out["probe_logits_sample"] = lg[:, ::s[0], ::s[1], ::s[2]].numpy()
        out["probe_logits_std"] = np.asarray(lg.std().item(), np.float32)
        out["meta"] = np.asarray([WEIGHT_SEED, seed, B], np.int64)
        np.savez_compressed(os.path.join(OUT, "fullnet_%s_%s.npz" % (robot, bb)), **out)
        print("fullnet", robot, bb, {k: tuple(v.shape) for k, v in out.items() if k.startswith(("joint", "kp2d_fk"))},
              "logit std %.3f" % lg.std().item())


def softargmax():
    ns = harness.setup()
    for path, nkpt, mode, B, seed in SA_CASES:
        hm = synth.make_heatmaps(B, nkpt, seed, mode)
        K, kv = synth.make_camera(B, seed)
        root = np.zeros((B, 3), np.float32)
        root[:, 2] = 0.8 + 0.1 * np.arange(B)
        layer = ns.integral.HeatmapIntegralPose(backbone=path, num_joints=nkpt, depth_dim=64, height_dim=64,
                                                width_dim=64, norm_type="softmax", image_size=256.0,
                                                bbox_3d_shape=[1300, 1300, 1300], rootid=3, fixroot=True)
        with torch.no_grad():
            uvd, xyz = layer(t(hm), root_trans=t(root), K=t(K))
        np.savez_compressed(os.path.join(OUT, "softargmax_%s_k%d_%s.npz" % (path, nkpt, mode)),
                            uvd=uvd.numpy(), xyz=xyz.numpy(), root_z=root[:, 2], meta=np.asarray([seed, B, nkpt]))
        print("softargmax", path, nkpt, mode, "uvd range", uvd.min().item(), uvd.max().item())


def fk():
    ns = harness.setup()
    for robot in ("panda", "kuka", "baxter"):
        r = ns.urdf_robot.URDFRobot(robot)
        assert [j.name for j in r.robot.actuated_joints] == ns.const.JOINT_NAMES[robot]
        q, rot, tr, K = synth.make_fk_inputs(robot, FK_N, FK_SEED)
        root = consts.ROBOTS[robot]["ref_kp"]
        with torch.no_grad():
            if root == 0:
                xyz = r.get_keypoints(t(q), t(rot), t(tr))
            else:
                xyz = r.get_keypoints_root(t(q), t(rot), t(tr), root=root)
            xyz0 = r.get_keypoints(t(q), t(rot), t(tr))      # base-rooted variant, exercised for every robot
            uv = ns.transforms.point_projection_from_3d_tensor(t(K), xyz)
        np.savez_compressed(os.path.join(OUT, "fk_%s.npz" % robot), xyz=xyz.numpy(), xyz_base_rooted=xyz0.numpy(),
                            uv=uv.numpy(), link_names=np.asarray(r.link_names),
                            offsets=r.offsets.numpy().reshape(-1, 3), meta=np.asarray([FK_SEED, FK_N, root]))
        print("fk", robot, xyz.shape, "link_names", list(r.link_names)[:4], "...")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    fk()
    softargmax()
    fullnet()
