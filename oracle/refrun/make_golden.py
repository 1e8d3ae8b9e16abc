"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only; see harness.py).

Fixtures (inputs are re-derivable from the recorded seeds through hrp_b200.synth, so only outputs are stored):
  fullnet_<robot>_<backbone>.npz  reference forward 8-tuple + caller-side projections on synthetic weights/inputs,
                                  plus stage-boundary probes captured with forward hooks (xf, img_feat, a strided
                                  sample of the heatmap logits);
  softargmax_<path>_k<nkpt>_<mode>.npz   HeatmapIntegralPose (integral.py:102-208) on adversarial heatmaps;
  fullnet_<robot>_<backbone>_undamped.npz   the same on synth's "undamped" weight recipe;
  fullnet_panda_resnet50_ckpt.npz the factory path get_rootNetwithRegInt_model with `pretrained_rootnet` (full_net.py:470-505:
                                  torch.load, `backbone.` -> `rootnet_backbone.` re-key, strict=False) followed by the evaluator's
                                  checkpoint load (fullnet_test.py:186-198), and a forward with init_pose / init_rot overrides;
  preprocess.npz                  the input side (8f N2): resize_image + CropResizeToAspectAugmentation + bbox_transform + k_value;
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
           ("panda", "hrnet32", 2, 2027), ("baxter", "hrnet32", 1, 2028), ("kuka", "hrnet32", 1, 2029)]
# recipe="undamped" weights (SURVEY 8d as written): reported per family, gated for fp32 only
UNDAMPED = [("panda", "resnet50", 2, 2031), ("panda", "hrnet32", 1, 2032)]
LOGIT_STRIDES = (37, 5, 7)
FK_N, FK_SEED = 512, 99
METRICS_N, METRICS_SEED = 400, 2718
JOINTMAP_GAIN = 6.0
SA_CASES = [("resnet50", 7, "blobs", 2, 11), ("resnet50", 7, "extreme", 2, 12), ("resnet50", 17, "noise", 1, 13),
            ("hrnet32", 8, "blobs", 2, 14), ("hrnet32", 17, "extreme", 1, 15)]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def fullnet(cases=None, recipe="damped", only=None):
    for robot, bb, B, seed in (cases or FULLNET):
        name = "fullnet_%s_%s%s.npz" % (robot, bb, "" if recipe == "damped" else "_" + recipe)
        if only and name not in only:
            continue
        sd = synth.make_state_dict(robot, bb, WEIGHT_SEED, recipe=recipe)
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
        np.savez_compressed(os.path.join(OUT, name), **out)
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


def checkpoint():
    """Factory + pretrained DepthNet re-key + evaluator-style checkpoint load, through the reference's own code."""
    import collections
    import tempfile
    ns = harness.setup()
    robot, bb, B, seed = "panda", "resnet50", 2, 2033
    sd = synth.make_state_dict(robot, bb, WEIGHT_SEED, recipe="undamped")
    pre = synth.make_pretrained_rootnet_state(sd)
    tmp = tempfile.mkdtemp(prefix="hrp_ckpt_")
    pre_path = os.path.join(tmp, "depthnet.pk")
    torch.save({"epoch": 3, "model_state_dict": collections.OrderedDict((k, t(v)) for k, v in pre.items())}, pre_path)
    cfg = harness.make_cfg(robot, bb)
    cfg.pretrained_rootnet = pre_path
    init = {"robot_type": robot, "pose_params": ns.const.INITIAL_JOINT_ANGLE, "cam_params": np.eye(4, dtype=float),
            "init_pose_from_mean": True}
    model = ns.full_net.get_rootNetwithRegInt_model(init, cfg)            # full_net.py:470-505
    # the evaluator then loads its checkpoint (DataParallel-prefixed keys, fullnet_test.py:186-198); here a checkpoint
    # of the keypoint branch + heads only, so the pretrained DepthNet stays in place
    main = collections.OrderedDict(("module." + k, t(v)) for k, v in sd.items()
                                   if not k.startswith(("rootnet_backbone.", "depth_layer.")))
    new_sd = collections.OrderedDict((k.replace("module.", "") if k.startswith("module.") else k, v) for k, v in main.items())
    model.load_state_dict(new_sd, strict=False)
    model.eval()
    img, K, kv = synth.make_inputs(B, seed)
    res = harness.forward(model, t(img), t(img), t(kv), t(K))
    out = {k: v.numpy() for k, v in res.items()}
    # init_pose / init_rot overrides (full_net.py:262, 268-272)
    g = np.random.Generator(np.random.PCG64(5))
    ip = (np.asarray(consts.ROBOTS[robot]["init_pose"], np.float32)[None] + 0.3 * g.standard_normal((B, 8))).astype(np.float32)
    ir = (np.asarray(consts.INIT_ROT6D, np.float32)[None] + 0.2 * g.standard_normal((B, 6))).astype(np.float32)
    with torch.no_grad():
        o2 = model(t(img), t(img), t(kv), t(K), init_pose=t(ip), init_rot=t(ir))
    out["init_pose"], out["init_rot"] = ip, ir
    out["ovr_joint_angles"], out["ovr_rot6d"], out["ovr_kp3d_fk"] = o2[0].numpy(), o2[1].numpy(), o2[7].numpy()
    out["meta"] = np.asarray([WEIGHT_SEED, seed, B], np.int64)
    np.savez_compressed(os.path.join(OUT, "fullnet_panda_resnet50_ckpt.npz"), **out)
    print("checkpoint", {k: tuple(v.shape) for k, v in out.items() if k.startswith(("joint", "ovr"))},
          "depth", out["root_depth"].ravel())


def preprocess():
    """The reference's own data-preparation functions on synthetic frames: lib/dataset/roboutils.py resize_image +
    bbox_transform, lib/dataset/augmentations.py CropResizeToAspectAugmentation (-> get_K_crop_resize), then the k_value
    expression of lib/core/function.py:98-110, strung together as lib/dataset/dream.py:415-449 does."""
    import copy
    harness.setup()
    from dataset import roboutils, augmentations
    n, seed = 6, 41
    frames, crop, kbox, K = synth.make_frames(n, seed)
    crops, Ks, kvs = [], [], []
    for i in range(n):
        state = {"camera": {"K": np.array(K[i], np.float64), "resolution": (640, 480)},
                 "objects": [{"keypoints_2d": np.zeros((1, 3)), "TCO_keypoints_3d": np.array([[0.1, 0.1, 1.0]]), "bbox": crop[i].copy()}]}
        K_original = copy.deepcopy(state["camera"]["K"])
        mask = np.zeros(frames[i].shape[:2], np.uint8)
        rgb, mask, state = roboutils.resize_image(np.asarray(frames[i]), crop[i], mask, state)
        m2 = np.ones(rgb.shape[:2], np.uint8)
        rgb, _, state = augmentations.CropResizeToAspectAugmentation(resize=(256, 256))(rgb, m2, state)
        rgb = augmentations.to_torch_uint8(rgb).permute(2, 0, 1)
        K_r = np.asarray(state["camera"]["K"])
        t = roboutils.bbox_transform(kbox[i], np.linalg.inv(K_original), K_r, resize_hw=(256, 256))
        t = torch.FloatTensor(np.array([max(0, t[0]), max(0, t[1]), min(256, t[2]), min(256, t[3])]))
        root_K = torch.FloatTensor(K_r)
        fx, fy = root_K[0, 0], root_K[1, 1]
        real_bbox = torch.tensor([1000.0, 1000.0]).to(torch.float32)
        area = torch.max(torch.abs(t[2] - t[0]), torch.abs(t[3] - t[1])) ** 2
        kv = torch.sqrt(fx * fy * real_bbox[0] * real_bbox[1] / area)
        crops.append(rgb.numpy()); Ks.append(root_K.numpy()); kvs.append(float(kv))
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), crops=np.stack(crops), K=np.stack(Ks), k_value=np.asarray(kvs, np.float32),
                        meta=np.asarray([seed, n], np.int64))
    print("preprocess", np.stack(crops).shape, "k_value", kvs)


# constructor variants outside the shipped configuration (full_net.py:107-131, 149-164): name -> reference cfg overrides
VARIANTS = {
    "direct_addfc_multikp": dict(direct_reg_rot=True, add_fc=True, multi_kp=True, kps_need_depth=[0, 3, 6]),
    "rotmatmul": dict(rot_iterative_matmul=True),
    "jointmap": dict(reg_joint_map=True, joint_conv_dim=[128, 64, 32]),
}


def variant_ctor(over, ref_kp):
    """The same switches in the form synth / the oracle / the CUDA path take."""
    kps = over.get("kps_need_depth") if over.get("multi_kp") else None
    return dict(direct_reg_rot=bool(over.get("direct_reg_rot", False)), rot_iterative_matmul=bool(over.get("rot_iterative_matmul", False)),
                add_fc=bool(over.get("add_fc", False)), depth_num=len(kps) if kps else 1, depth_root=kps.index(ref_kp) if kps else 0,
                reg_joint_map=bool(over.get("reg_joint_map", False)), joint_conv_dim=tuple(over.get("joint_conv_dim", ())))


def variants():
    robot, bb, B, seed = "panda", "resnet50", 3, 31
    for name, over in VARIANTS.items():
        ctor = variant_ctor(over, consts.ROBOTS[robot]["ref_kp"])
        sd = synth.make_state_dict(robot, bb, WEIGHT_SEED, ctor={k: v for k, v in ctor.items() if k != "depth_root"})
        if ctor["reg_joint_map"]:            # as drawn, the joint maps are nearly flat (every angle mid-range): give them contrast
            sd["joint_final_layer.weight"] = sd["joint_final_layer.weight"] * np.float32(JOINTMAP_GAIN)
        model, _ = harness.build_model(robot, bb, **over)
        model.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
        img, K, kv = synth.make_inputs(B, seed)
        res = harness.forward(model, t(img), t(img), t(kv), t(K))
        out = {k: v.numpy() for k, v in res.items()}
        out["meta"] = np.asarray([WEIGHT_SEED, seed, B], np.int64)
        np.savez_compressed(os.path.join(OUT, "variant_%s.npz" % name), **out)
        print("variant", name, ctor, "rot6d[0]", out["rot6d"][0].round(3), "depth", out["root_depth"].ravel().round(3))


def metrics():
    """compute_metrics_batch + summary_add_pck of the reference itself (lib/utils/metrics.py) on synth.make_metrics_inputs."""
    ns = harness.setup()
    import warnings
    from utils import metrics as ref_metrics
    for robot in ("panda", "kuka", "baxter"):
        r = ns.urdf_robot.URDFRobot(robot)
        root = consts.ROBOTS[robot]["ref_kp"]
        d = synth.make_metrics_inputs(robot, METRICS_N, METRICS_SEED)
        with torch.no_grad():
            kp = (lambda q, ro, tr: r.get_keypoints(q, ro, tr)) if root == 0 else (lambda q, ro, tr: r.get_keypoints_root(q, ro, tr, root=root))
            gt_xyz = kp(t(d["gt_q"]), t(d["gt_rot"]), t(d["gt_trans"]))
            gt_uv = ns.transforms.point_projection_from_3d_tensor(t(d["K"]), gt_xyz)
            gt_uv[3] = -50.0                                  # one frame with no keypoint inside the image: error2d = 0/0
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = ref_metrics.compute_metrics_batch(r, gt_xyz, gt_uv, t(d["K"]), t(d["gt_q"]), pred_joint=t(d["q"]), pred_rot=t(d["rot"]),
                                                        pred_trans=t(d["trans"]), pred_depth=None, pred_xy=None, pred_xyz_integral=None,
                                                        reference_keypoint_id=root)
                pred_xyz = kp(t(d["q"]), t(d["rot"]), t(d["trans"]))
                res_nojoint = ref_metrics.compute_metrics_batch(r, gt_xyz, gt_uv, t(d["K"]), t(d["gt_q"]), pred_joint=None, pred_rot=None, pred_trans=None,
                                                                pred_xyz_integral=pred_xyz, reference_keypoint_id=root)
                ok = ~np.isnan(np.asarray(res[1]))
                alldis = {"dis3d": list(np.asarray(res[0])[ok]), "dis2d": list(np.asarray(res[1])[ok])}
                summ = ref_metrics.summary_add_pck(alldis)
        names = ["error3d", "error2d", "dis3d", "dis2d", "l1_jointerror", "mean_jointerror", "error_depth", "batch_error_relative", "error3d_relative"]
        out = {n_: np.asarray(v) for n_, v in zip(names, res)}
        out.update({"nojoint_" + n_: np.asarray(v) for n_, v in zip(names, res_nojoint)})
        out["summary_keys"] = np.asarray(list(summ.keys()))
        out["summary_values"] = np.asarray([float(v) for v in summ.values()], np.float64)
        out["summary_dtypes"] = np.asarray([type(v).__name__ for v in summ.values()])
        out["gt_xyz"] = gt_xyz.numpy(); out["gt_uv"] = gt_uv.numpy(); out["pred_xyz"] = pred_xyz.numpy()
        out["meta"] = np.asarray([METRICS_SEED, METRICS_N, root])
        np.savez_compressed(os.path.join(OUT, "metrics_%s.npz" % robot), **out)
        print("metrics", robot, {k: (round(float(v), 5)) for k, v in list(summ.items())[:6]}, "error2d NaN frames:", int((~ok).sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    what = sys.argv[1:] or ["fk", "softargmax", "fullnet", "undamped", "checkpoint", "preprocess", "metrics", "variants"]
    if "fk" in what:
        fk()
    if "softargmax" in what:
        softargmax()
    if "fullnet" in what:
        fullnet()
    if "fullnet_new" in what:
        fullnet(only={"fullnet_kuka_hrnet32.npz"})
    if "undamped" in what:
        fullnet(UNDAMPED, recipe="undamped")
    if "checkpoint" in what:
        checkpoint()
    if "preprocess" in what:
        preprocess()
    if "metrics" in what:
        metrics()
    if "variants" in what:
        variants()
