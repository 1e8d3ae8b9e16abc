#!/usr/bin/env python
"""A small forward of every code path (families, constructor variants, input side, evaluation tail) for compute-sanitizer:
compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hrp_b200  # noqa
from hrp_b200 import synth, metrics as hm
from hrp_b200.model import HoliRobPoseB200, FkRobot, crop_resize
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, 3))
for prec in ("f16", "tf32", "fp32"):
    m = HoliRobPoseB200("panda", device=dev, precision=prec)
    m.load_state_dict(synth.make_state_dict("panda", "resnet50"))
    out = m.forward_dict(img, K, kv)
    torch.cuda.synchronize()
    print(prec, "ok", float(out["joint_angles"].abs().max()))
    del m
for cfg, ctor in ((dict(direct_reg_rot=True, add_fc=True, multi_kp=True, kps_need_depth=[0, 3, 6]), dict(direct_reg_rot=True, add_fc=True, depth_num=3)),
                  (dict(rot_iterative_matmul=True), dict(rot_iterative_matmul=True)),
                  (dict(reg_joint_map=True, joint_conv_dim=[128, 64, 32]), dict(reg_joint_map=True, joint_conv_dim=(128, 64, 32)))):
    m = HoliRobPoseB200("panda", cfg, device=dev, precision="f16")
    m.load_state_dict(synth.make_state_dict("panda", "resnet50", ctor=ctor))
    out = m.forward_dict(img, K, kv)
    torch.cuda.synchronize()
    print(sorted(cfg), "ok")
    del m
frames, crop, kbox, Kf = synth.make_frames(2, 5)
c, Kc, kvc = crop_resize(torch.from_numpy(frames).to(dev), torch.from_numpy(np.asarray(crop, np.int32)).to(dev), torch.from_numpy(Kf).to(dev),
                         torch.from_numpy(np.asarray(kbox, np.float32)).to(dev))
torch.cuda.synchronize()
print("crop_resize ok", tuple(c.shape))
d = synth.make_metrics_inputs("panda", 100, 1)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
fk = FkRobot("panda")
gxyz, guv = fk.keypoints(T(d["gt_q"]), T(d["gt_rot"]), T(d["gt_trans"]), T(d["K"]))
pf, d3, d2, l1 = hm.metrics_batch_device(fk, gxyz, guv, T(d["K"]), T(d["gt_q"]), pred_joint=T(d["q"]), pred_rot=T(d["rot"]), pred_trans=T(d["trans"]))
s = hm.summary_add_pck({"dis3d": pf[:, 0].contiguous(), "dis2d": torch.nan_to_num(pf[:, 1], nan=1e3).contiguous()})
torch.cuda.synchronize()
print("metrics ok", float(s["ADD/AUC"]))
