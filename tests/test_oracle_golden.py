"""CPU: the oracle port reproduces the reference's own outputs (tests/golden/*, made by oracle/refrun/make_golden.py)."""
import numpy as np
import pytest
import torch

from hrp_b200 import consts, synth
from oracle import integral, kinematics

import helpers


@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_fk_matches_reference(robot):
    g = helpers.load_golden("fk_%s.npz" % robot)
    seed, n, root = (int(v) for v in g["meta"])
    q, rot, tr, K = synth.make_fk_inputs(robot, n, seed)
    om, _ = helpers.oracle_for(robot, "resnet50")
    assert om.actuated_names == consts.ROBOTS[robot]["joints"]
    assert [l for l, _ in om.kp_frames] == [str(s) for s in g["link_names"]]
    np.testing.assert_allclose(np.stack([o for _, o in om.kp_frames]), g["offsets"], atol=1e-7)
    xyz = kinematics.keypoints(om.kin, om.kp_frames, q, rot, tr, root)
    assert helpers.maxdiff(xyz, g["xyz"]) < 2e-6
    xyz0 = kinematics.keypoints(om.kin, om.kp_frames, q, rot, tr, 0)
    assert helpers.maxdiff(xyz0, g["xyz_base_rooted"]) < 2e-6
    uv = integral.project(torch.from_numpy(K), torch.from_numpy(xyz)).numpy()
    ok = np.abs(g["xyz"][..., 2]) > 0.05          # projection is ill-conditioned as z -> 0
    assert np.max(np.abs(uv - g["uv"])[ok]) < 2e-2


@pytest.mark.parametrize("robot", ["panda", "kuka"])
def test_fk_limb_lengths_known_answer(robot):
    """In-tree known answers: lib/dataset/const.py:108-124."""
    q, rot, tr, _ = synth.make_fk_inputs(robot, 64, 5)
    om, _ = helpers.oracle_for(robot, "resnet50")
    xyz = kinematics.keypoints(om.kin, om.kp_frames, q, rot, tr, om.ref)
    d = np.linalg.norm(xyz[:, 1:] - xyz[:, :-1], axis=2)
    np.testing.assert_allclose(d, np.broadcast_to(consts.LIMB_LENGTH[robot], d.shape), atol=2e-5)


def test_rot6d_identity_known_answer():
    """rot6d(I) = [1,0,0,0,1,0] (full_net.py:205) maps back to the identity."""
    R = kinematics.rot6d_to_rotmat(np.asarray([consts.INIT_ROT6D], np.float32))
    np.testing.assert_allclose(R[0], np.eye(3), atol=0)


@pytest.mark.parametrize("case", ["resnet50_k7_blobs", "resnet50_k7_extreme", "resnet50_k17_noise",
                                  "hrnet32_k8_blobs", "hrnet32_k17_extreme"])
def test_softargmax_matches_reference(case):
    g = helpers.load_golden("softargmax_%s.npz" % case)
    seed, B, nkpt = (int(v) for v in g["meta"])
    mode = case.split("_")[-1]
    hm = torch.from_numpy(synth.make_heatmaps(B, nkpt, seed, mode))
    K, _ = synth.make_camera(B, seed)
    uvd = integral.soft_argmax_uvd(hm, nkpt, 3, True, path=case.split("_")[0])
    xyz = integral.uvd_to_xyz(uvd, torch.from_numpy(K), torch.from_numpy(g["root_z"]), 256.0, 1.3)
    assert helpers.maxdiff(uvd, g["uvd"]) < 2e-6
    assert helpers.maxdiff(xyz, g["xyz"]) < 5e-6


@pytest.mark.parametrize("robot,backbone", helpers.FULLNET_CASES)
def test_fullnet_matches_reference(robot, backbone):
    g = helpers.load_golden("fullnet_%s_%s.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    om, _ = helpers.oracle_for(robot, backbone, wseed)
    img, K, kv = helpers.inputs(B, seed)
    trace = {}
    out = om.forward(img, img, kv, K, trace=trace)
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk"]
    res = dict(zip(names, out))
    res["kp2d_int"] = integral.project(K, res["kp3d_int"])
    res["kp2d_fk"] = integral.project(K, res["kp3d_fk"])
    tol = dict(joint_angles=1e-5, rot6d=1e-5, trans=1e-5, root_uv=1e-3, root_depth=1e-5, uvd=1e-5, kp3d_int=1e-5,
               kp3d_fk=1e-5, kp2d_int=5e-3, kp2d_fk=5e-3)
    for k, t in tol.items():
        assert tuple(res[k].shape) == g[k].shape, k
        assert helpers.maxdiff(res[k], g[k]) < t, (k, helpers.maxdiff(res[k], g[k]))
    s = (37, 5, 7)
    assert helpers.maxdiff(trace["xf"], g["probe_xf"]) < 1e-4
    assert helpers.maxdiff(trace["img_feat"], g["probe_img_feat"]) < 1e-4
    assert helpers.maxdiff(trace["logits"][:, ::s[0], ::s[1], ::s[2]], g["probe_logits_sample"]) < 1e-3
    # the fixture is non-trivial: peaked-but-not-one-hot heatmaps, metre-scale depths, radians-scale angles
    assert 1.0 < float(g["probe_logits_std"]) < 5.0
    assert 0.3 < float(np.abs(g["root_depth"]).max()) < 10.0


@pytest.mark.parametrize("robot,backbone", helpers.UNDAMPED_CASES)
def test_fullnet_undamped_weights_match_reference(robot, backbone):
    """The SURVEY 8d weight recipe as written (unit-gain residual branches, its own BN calibration)."""
    g = helpers.load_golden("fullnet_%s_%s_undamped.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    om, _ = helpers.oracle_for(robot, backbone, wseed, recipe="undamped")
    img, K, kv = helpers.inputs(B, seed)
    res = om.forward_dict(img, img, kv, K)
    for k, t in dict(joint_angles=2e-5, rot6d=2e-5, root_depth=2e-5, uvd=2e-5, kp3d_fk=2e-5, kp2d_int=1e-2, kp2d_fk=1e-2).items():
        assert helpers.maxdiff(res[k], g[k]) < t, (k, helpers.maxdiff(res[k], g[k]))


def test_checkpoint_merge_and_init_overrides_match_reference():
    """Golden made by the reference's factory with `pretrained_rootnet` (torch.load + `backbone.` -> `rootnet_backbone.`
    re-key + strict=False, full_net.py:486-500) followed by the evaluator's `module.`-stripping load
    (fullnet_test.py:186-198); plus a forward with init_pose / init_rot (full_net.py:268-272)."""
    g = helpers.load_golden("fullnet_panda_resnet50_ckpt.npz")
    wseed, seed, B = (int(v) for v in g["meta"])
    _, _, merged = helpers.checkpoint_case()
    from oracle import model as omodel
    om = omodel.OracleModel("panda", merged, open(consts.urdf_path("panda")).read(), "resnet50")
    img, K, kv = helpers.inputs(B, seed)
    res = om.forward_dict(img, img, kv, K)
    for k, t in dict(joint_angles=2e-5, rot6d=2e-5, root_depth=2e-5, kp3d_fk=2e-5, kp2d_fk=1e-2).items():
        assert helpers.maxdiff(res[k], g[k]) < t, (k, helpers.maxdiff(res[k], g[k]))
    trace = {}
    o2 = om.forward(img, img, kv, K, trace=trace, init_pose=torch.from_numpy(g["init_pose"]), init_rot=torch.from_numpy(g["init_rot"]))
    assert helpers.maxdiff(o2[0], g["ovr_joint_angles"]) < 2e-5 and helpers.maxdiff(o2[1], g["ovr_rot6d"]) < 2e-5
    assert helpers.maxdiff(o2[7], g["ovr_kp3d_fk"]) < 2e-5
    assert helpers.maxdiff(o2[0], res["joint_angles"]) > 1e-2          # the override matters
    assert len(trace["pose_iters"]) == 4


def test_preprocess_matches_reference():
    """Input side (8f N2): the port of resize_image + CropResizeToAspectAugmentation + get_K_crop_resize + bbox_transform +
    k_value against what the reference's own functions produced (bit-exact crops: same torch CPU interpolate)."""
    from oracle import preprocess
    g = helpers.load_golden("preprocess.npz")
    seed, n = (int(v) for v in g["meta"])
    frames, crop, kbox, K = synth.make_frames(n, seed)
    for i in range(n):
        c, Kn, kv = preprocess.crop_resize_one(frames[i], crop[i], K[i], kbox[i])
        assert np.array_equal(c, g["crops"][i]), i
        np.testing.assert_allclose(Kn, g["K"][i], rtol=0, atol=1e-4)
        np.testing.assert_allclose(kv, g["k_value"][i], rtol=1e-6)
    assert crop[0][2] - crop[0][0] == 256          # the no-resize branch is part of the fixture


@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_metrics_match_reference(robot):
    """Evaluation tail (8f N4): the port of compute_metrics_batch / summary_add_pck against what the reference's own functions
    returned (metrics.py:8-162), including a frame with no keypoint inside the image (error2d = NaN) and the no-joint branch."""
    from oracle import metrics as ometrics
    g = helpers.load_golden("metrics_%s.npz" % robot)
    seed, n, root = (int(v) for v in g["meta"])
    d = synth.make_metrics_inputs(robot, n, seed)
    names = ["error3d", "error2d", "dis3d", "dis2d", "l1_jointerror", "mean_jointerror", "error_depth", "batch_error_relative", "error3d_relative"]
    res = ometrics.batch_errors(g["pred_xyz"], g["gt_xyz"], g["gt_uv"], d["K"], d["gt_q"], d["q"], root, robot)
    for k, v in zip(names, res):
        np.testing.assert_allclose(np.asarray(v), g[k], rtol=2e-5, atol=2e-6, equal_nan=True, err_msg=k)
    res = ometrics.batch_errors(g["pred_xyz"], g["gt_xyz"], g["gt_uv"], d["K"], d["gt_q"], None, root, robot)
    for k, v in zip(names, res):
        np.testing.assert_allclose(np.asarray(v), g["nojoint_" + k], rtol=2e-5, atol=2e-6, equal_nan=True, err_msg=k)
    assert np.isnan(g["error2d"]).sum() >= 1 and np.isnan(g["error2d"][3])
    ok = ~np.isnan(g["error2d"])
    s = ometrics.summary(g["error3d"][ok], g["error2d"][ok])
    assert list(s.keys()) == [str(k) for k in g["summary_keys"]]
    for k, v in zip(g["summary_keys"], g["summary_values"]):
        np.testing.assert_allclose(float(s[str(k)]), v, rtol=1e-6, atol=1e-9, err_msg=str(k))


@pytest.mark.parametrize("name", list(helpers.VARIANT_CASES))
def test_constructor_variants_match_reference(name):
    """8f N4: direct_reg_rot + add_fc + multi_kp, rot_iterative_matmul, and reg_joint_map + HeatmapIntegralJoint
    (full_net.py:92-131, 149-164, 293-330, 376-429; integral.py:211-251), against the reference constructed with those switches."""
    g = helpers.load_golden("variant_%s.npz" % name)
    wseed, seed, B = (int(v) for v in g["meta"])
    from oracle import model as omodel
    om = omodel.OracleModel("panda", helpers.variant_state_dict(name, wseed), open(consts.urdf_path("panda")).read(), "resnet50",
                            ctor=helpers.VARIANT_CASES[name][1])
    img, K, kv = helpers.inputs(B, seed)
    res = om.forward_dict(img, img, kv, K)
    for k, t in dict(joint_angles=2e-5, rot6d=2e-5, trans=2e-5, root_depth=2e-5, uvd=2e-5, kp3d_fk=2e-5, kp2d_int=1e-2, kp2d_fk=1e-2).items():
        assert helpers.maxdiff(res[k], g[k]) < t, (k, helpers.maxdiff(res[k], g[k]))
    if "depths" in g:                                                              # multi_kp: the 9-tuple's extra entry
        trace = {}
        om.forward(img, img, kv, K, trace=trace)
        assert g["depths"].shape == (B, 3) and helpers.maxdiff(trace["depths"], g["depths"]) < 2e-5
        assert helpers.maxdiff(g["depths"][:, 1:2], g["root_depth"]) == 0.0
    shipped, _ = helpers.oracle_for("panda", "resnet50", wseed)
    base = shipped.forward_dict(img, img, kv, K)
    changed = "joint_angles" if name == "jointmap" else "rot6d"
    assert helpers.maxdiff(base[changed], g[changed]) > 1e-2                        # the switch matters
