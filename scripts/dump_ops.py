#!/usr/bin/env python
"""Per-op device times of one un-graphed forward (HRP_DUMP_OPS csv): python scripts/dump_ops.py out.csv [precision] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200
out = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda", 0)
m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision=prec)
m.load_state_dict(synth.make_state_dict("panda", "resnet50"))
img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, 1))
for _ in range(2):
    m.profile(img, img, kv, K)
os.environ["HRP_DUMP_OPS"] = out
m.profile(img, img, kv, K)
