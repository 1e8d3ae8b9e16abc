#!/usr/bin/env python
"""Three forwards of the bench configuration and nothing else (for an ncu launch list: the last N launches are one graph
replay): python scripts/one_step.py [precision] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision=prec)
m.load_state_dict(synth.make_state_dict("panda", "resnet50"))
m.set_option("slots", 1)
img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, 1))
for _ in range(3):
    out = m.forward_dict(img, K, kv)
torch.cuda.synchronize()
print("launches per forward:", m.launch_count())
