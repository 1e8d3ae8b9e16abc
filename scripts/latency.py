#!/usr/bin/env python
"""Single-call latency of the forward (BASELINE configs[0] is batch 1): latency.py [precision] [lane_share_pct, 0 = library default]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
share = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision=prec)
m.load_state_dict(synth.make_state_dict("panda", "resnet50"))
if share > 0:
    m.set_option("lane_share_pct", share)
for B in (1, 4, 16, 64):
    img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(B, 1))
    for _ in range(5):
        m.forward_dict(img, K, kv)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.forward_dict(img, K, kv); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("%s share %d%% batch %3d: median %.3f ms (min %.3f) -> %.0f frames/s one call at a time" % (prec, share, B, ts[10], ts[0], B / ts[10] * 1e3))
