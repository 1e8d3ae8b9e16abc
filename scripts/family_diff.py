#!/usr/bin/env python
"""Development probe: where do the tensor-core families depart from the fp32 family inside the network?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200

dev = torch.device("cuda", 0)
robot, backbone, B = "panda", sys.argv[1] if len(sys.argv) > 1 else "resnet50", 4
sd = synth.make_state_dict(robot, backbone)
img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, 99))
outs, dbg = {}, {}
for prec in ("fp32", "tf32", "bf16"):
    m = HoliRobPoseB200(robot, {"backbone_name": backbone}, device=dev, precision=prec)
    m.load_state_dict(sd)
    outs[prec] = {k: v.clone() for k, v in m.forward_dict(img, K, kv).items()}
    dbg[prec] = {n: m.debug_tensor(n, B).clone() for n in ("xf", "img_feat", "logits")}
for prec in ("tf32", "bf16"):
    print(prec, {k: "%.2e" % float((outs[prec][k] - outs["fp32"][k]).abs().max()) for k in outs["fp32"]})
    for n in dbg["fp32"]:
        a, b = dbg[prec][n], dbg["fp32"][n]
        print("   %-9s max|d| %.3e  rel-rms %.3e  (fp32 rms %.3e)" % (n, float((a - b).abs().max()), float((a - b).norm() / b.norm()), float(b.pow(2).mean().sqrt())))
