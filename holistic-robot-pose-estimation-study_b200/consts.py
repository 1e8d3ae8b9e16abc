"""Per-robot constants of the HoliRobPose inference path.

Data restated from the reference (facts, not code): robot table lib/models/full_net.py:42-53, keypoint links and
actuated-joint order lib/dataset/const.py:61-90, iteration seeds (`INITIAL_JOINT_ANGLE['mean']`) const.py:185-236,
joint bounds const.py:239-284, shipped `reference_keypoint_id` / `bbox_3d_shape` configs/{panda,kuka,baxter}/full.yaml.
"""
import os

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

DEPTH_DIM = 64          # full_net.py:66
HEATMAP_SIZE = 64       # image_size / 4, full_net.py:67-68
IMAGE_SIZE = 256        # other_image_size, configs/*/full.yaml
FEATURE_DIM = 2048      # full_net.py:79,86,149
N_ITER = 4              # configs/*/full.yaml n_iter
ROT_DIM = 6             # lib/core/config.py rotation_dim default

ROBOTS = {
    "panda": dict(
        dof=8, nkpt=7, ref_kp=3, bbox_3d=(1300.0, 1300.0, 1300.0), urdf="panda.urdf",
        links=["panda_link0", "panda_link2", "panda_link3", "panda_link4", "panda_link6", "panda_link7", "panda_hand"],
        joints=["panda_joint1", "panda_joint2", "panda_joint3", "panda_joint4", "panda_joint5", "panda_joint6",
                "panda_joint7", "panda_finger_joint1"],
        init_pose=[0.0, 0.0, 0.0, -1.52715, 0.0, 1.8675, 0.0, 0.02],
        bounds=[[-2.9671, 2.9671], [-1.8326, 1.8326], [-2.9671, 2.9671], [-3.1416, 0.0873], [-2.9671, 2.9671],
                [-0.0873, 3.8223], [-2.9671, 2.9671], [0.0, 0.04]],
    ),
    "kuka": dict(
        dof=7, nkpt=8, ref_kp=3, bbox_3d=(1300.0, 1300.0, 1300.0), urdf="iiwa7.urdf",
        links=["iiwa_link_%d" % i for i in range(8)],
        joints=["iiwa_joint_%d" % i for i in range(1, 8)],
        init_pose=[0.0] * 7,
        bounds=[[-2.9671, 2.9671], [-2.0944, 2.0944], [-2.9671, 2.9671], [-2.0944, 2.0944], [-2.9671, 2.9671],
                [-2.0944, 2.0944], [-3.0543, 3.0543]],
    ),
    "baxter": dict(
        dof=15, nkpt=17, ref_kp=0, bbox_3d=(1300.0, 1300.0, 1300.0), urdf="baxter.urdf",
        # Baxter keypoints are the origins of these joints expressed in their parent link (urdf_robot.py:68-87);
        # the link list is derived from the URDF at load time.
        kp_joints=["torso_t0", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1",
                   "left_e1", "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2", "right_hand",
                   "left_hand"],
        joints=["head_pan", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1", "left_e1",
                "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2"],
        init_pose=[0.0, 0.0, 0.0, -0.5499999999999999, -0.5499999999999999, 0.0, 0.0, 1.284, 1.284, 0.0, 0.0,
                   0.2616018366049999, 0.2616018366049999, 0.0, 0.0],
        bounds=[[-1.5708, 1.5708], [-1.7017, 1.7017], [-1.7017, 1.7017], [-2.1470, 1.0470], [-2.1470, 1.0470],
                [-3.0542, 3.0542], [-3.0542, 3.0542], [-0.0500, 2.6180], [-0.0500, 2.6180], [-3.0590, 3.0590],
                [-3.0590, 3.0590], [-1.5708, 2.0940], [-1.5708, 2.0940], [-3.0590, 3.0590], [-3.0590, 3.0590]],
    ),
}

# Known-answer limb lengths (const.py:108-124): |p(link_a) - p(link_b)| for consecutive keypoint links, any pose.
LIMB_LENGTH = {
    "panda": [0.3330, 0.3160, 0.0825, 0.39276, 0.0880, 0.1070],
    "kuka": [0.1500, 0.1900, 0.2100, 0.1900, 0.2100, 0.19946, 0.10122],
}

INIT_ROT6D = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]  # rotmat_to_rot6d(I), full_net.py:205, geometries.py:117-132


def urdf_path(robot):
    return os.path.join(DATA_DIR, "urdf", ROBOTS[robot]["urdf"])
