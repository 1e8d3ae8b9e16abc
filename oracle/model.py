"""Oracle: the full inference forward `model(x_reg, x_root, k_value, K)` + caller-side projections.

Follows lib/models/full_net.py:262-466 (shipped configuration: n_iter=4, rotation_dim=6, fix_root=True; the
constructor variants direct_reg_rot / rot_iterative_matmul / add_fc / multi_kp / reg_joint_map through `ctor`) and lib/core/function.py:133-141.
"""
import numpy as np
import torch

from . import integral, kinematics, network

ROBOTS = {  # (dof, nkpt, reference_keypoint_id, urdf file) -- full_net.py:42-53, configs/*/full.yaml
    "panda": (8, 7, 3, "panda.urdf"),
    "kuka": (7, 8, 3, "iiwa7.urdf"),
    "baxter": (15, 17, 0, "baxter.urdf"),
}
LINKS = {  # lib/dataset/const.py:61-69
    "panda": ["panda_link0", "panda_link2", "panda_link3", "panda_link4", "panda_link6", "panda_link7", "panda_hand"],
    "kuka": ["iiwa_link_%d" % i for i in range(8)],
}
BAXTER_KP_JOINTS = ["torso_t0", "right_s0", "left_s0", "right_s1", "left_s1", "right_e0", "left_e0", "right_e1",
                    "left_e1", "right_w0", "left_w0", "right_w1", "left_w1", "right_w2", "left_w2", "right_hand",
                    "left_hand"]  # urdf_robot.py:73-77


class OracleModel:
    def __init__(self, robot, state_dict, urdf_text, backbone="resnet50", image_size=256.0, depth_mm=1300.0,
                 fix_root=True, n_iter=4, ctor=None):
        # constructor switches outside the shipped configuration (full_net.py:107-131, 149-164, 293-330, 395-444):
        # direct_reg_rot, rot_iterative_matmul, add_fc, depth_num (= len(kps_need_depth)) + depth_root (index of the root keypoint in it)
        self.ctor = dict(direct_reg_rot=False, rot_iterative_matmul=False, add_fc=False, depth_num=1, depth_root=0, reg_joint_map=False,
                         joint_bounds=None)     # reg_joint_map (resnet50 only): joint angles from joint_map_head; bounds const.py:239-284
        self.ctor.update(ctor or {})
        self.robot = robot
        self.dof, self.nkpt, self.ref, _ = ROBOTS[robot]
        self.backbone = "resnet50" if backbone in ("resnet", "resnet50") else "hrnet32"
        self.sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)))
                   for k, v in state_dict.items()}
        self.image_size = image_size
        self.depth_factor = float(np.float32(depth_mm) * np.float32(1e-3))   # integral.py:96-97
        self.fix_root = fix_root
        self.n_iter = n_iter
        self.kin = kinematics.OracleRobot(urdf_text)
        names = [j["name"] for j in self.kin.actuated]
        if robot == "baxter":
            self.kp_frames = [(self.kin.by_name[j]["parent"], self.kin.by_name[j]["origin"][:3, 3])
                              for j in BAXTER_KP_JOINTS]
        else:
            self.kp_frames = [(l, np.zeros(3)) for l in LINKS[robot]]
        self.actuated_names = names

    def fk(self, pose, rot, trans):
        return kinematics.keypoints(self.kin, self.kp_frames, np.asarray(pose, np.float32), np.asarray(rot, np.float32),
                                    np.asarray(trans, np.float32), self.ref)

    @torch.no_grad()
    def forward(self, x_reg, x_root, k_value, K, calib=None, trace=None, init_pose=None, init_rot=None):
        sd = self.sd
        x_reg, x_root = x_reg.float(), x_root.float()
        B = x_reg.shape[0]
        _, img_feat = network.hrnet_w32(x_root, sd, "rootnet_backbone.", False, calib)
        feat = network.depth_add_fc(img_feat, sd) if self.ctor["add_fc"] else img_feat
        gamma = torch.nn.functional.conv2d(feat[:, :, None, None], sd["depth_layer.weight"], sd["depth_layer.bias"])
        if self.ctor["depth_num"] > 1:                                                      # multi_kp, full_net.py:319-329
            depths = gamma.view(-1, self.ctor["depth_num"]) * k_value.view(-1, 1).expand(-1, self.ctor["depth_num"]) / 1000.0
            depth = depths[:, self.ctor["depth_root"]].reshape(-1, 1)
            if trace is not None:
                trace["depths"] = depths
        else:
            depth = (gamma.view(-1, 1) * k_value.view(-1, 1)).reshape(B, 1) / 1000.0      # full_net.py:334-336
        if self.backbone == "resnet50":
            x_out = network.resnet50(x_reg, sd, "reg_backbone.", calib)
            xf = torch.nn.functional.avg_pool2d(x_out, 8, 1).view(B, -1)
            logits = network.deconv_head(x_out, sd, calib)
        else:
            logits, xf = network.hrnet_w32(x_reg, sd, "reg_backbone.", True, calib)
        uvd = integral.soft_argmax_uvd(logits, self.nkpt, self.ref, self.fix_root, path=self.backbone)
        xyz_int = integral.uvd_to_xyz(uvd, K, depth[:, 0], self.image_size, self.depth_factor)
        root_uv = (uvd[:, self.ref, :2] + 0.5) * self.image_size
        trans = integral.uvz_to_xyz(root_uv, depth, K)
        tp = [] if trace is not None else None
        tr = [] if trace is not None else None
        p0 = sd["init_pose"].expand(B, -1) if init_pose is None else init_pose        # full_net.py:268-272
        r0 = sd["init_rot"].expand(B, -1) if init_rot is None else init_rot
        if self.ctor["reg_joint_map"]:
            pose = network.joint_map_head(x_out, sd, self.ctor["joint_bounds"])
        else:
            pose = network.iterative_head(xf, p0, sd, "fc_pose_1", "fc_pose_2", "decpose", self.n_iter, tp)
        if self.ctor["direct_reg_rot"]:
            rot = network.direct_rot_head(xf, sd)
        elif self.ctor["rot_iterative_matmul"]:
            rot = network.matmul_rot_head(xf, r0, sd, self.n_iter, tr)
        else:
            rot = network.iterative_head(xf, r0, sd, "fc_rot_1", "fc_rot_2", "decrot", self.n_iter, tr)
        xyz_fk = torch.from_numpy(self.fk(pose.numpy(), rot.numpy(), trans.numpy()))
        if trace is not None:
            trace.update(logits=logits, xf=xf, img_feat=img_feat, pose_iters=tp, rot_iters=tr)
        return pose, rot, trans, root_uv, depth, uvd, xyz_int, xyz_fk

    def forward_dict(self, x_reg, x_root, k_value, K):
        out = self.forward(x_reg, x_root, k_value, K)
        names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk"]
        res = {n: o for n, o in zip(names, out)}
        res["kp2d_int"] = integral.project(K, res["kp3d_int"])
        res["kp2d_fk"] = integral.project(K, res["kp3d_fk"])
        return res
