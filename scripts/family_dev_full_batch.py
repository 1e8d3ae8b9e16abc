#!/usr/bin/env python
"""Deviation of each precision family from the oracle over EVERY frame of a full batch (not a sample):
python scripts/family_dev_full_batch.py [robot] [batch] -> one JSON line per family (max and 99th percentile per field)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hrp_b200  # noqa
from hrp_b200 import consts, synth
from hrp_b200.model import HoliRobPoseB200
from oracle import model as omodel          # checker
robot = sys.argv[1] if len(sys.argv) > 1 else "panda"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
sd = synth.make_state_dict(robot, "resnet50")
img, K, kv = (torch.from_numpy(a) for a in synth.make_inputs(B, 4000 + B))
torch.set_num_threads(len(os.sched_getaffinity(0)))
ref = omodel.OracleModel(robot, sd, open(consts.urdf_path(robot)).read(), "resnet50").forward_dict(img, img, kv, K)
for prec in ("fp32", "tf32", "f16", "bf16"):
    m = HoliRobPoseB200(robot, {"backbone_name": "resnet50"}, device=dev, precision=prec)
    m.load_state_dict(sd)
    out = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    row = {"robot": robot, "batch": B, "precision": prec}
    for k in ("joint_angles", "root_depth", "root_uv", "kp2d_int", "kp2d_fk"):
        d = (out[k].cpu() - ref[k]).abs().reshape(B, -1).max(1).values.numpy()
        row[k] = {"max": float(d.max()), "p99": float(np.percentile(d, 99)), "median": float(np.median(d)), "frames_over_gate": int((d > {"joint_angles": 1e-3, "root_depth": 1e-3}.get(k, 0.5)).sum())}
    print(json.dumps(row))
    del m
