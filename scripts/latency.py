#!/usr/bin/env python
"""Single-call latency table of the forward, all rows in ONE run (BASELINE configs[0] is batch 1):
python scripts/latency.py            one call at a time, CUDA events around forward_dict on resident inputs, median of 20."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200
dev = torch.device("cuda", 0)
sd = synth.make_state_dict("panda", "resnet50")
rows = [("bf16", 0, (1, 4, 16, 64)), ("bf16", 50, (1, 16, 64)), ("bf16", 100, (64,)), ("tf32", 0, (1, 64)), ("tf32", 100, (64,)), ("fp32", 0, (1,))]
print("Single-call latency of the full forward (Panda, shipped configuration, B200): one call at a time, CUDA events around\n"
      "forward_dict on resident fp32 inputs, median of 20 after 5 warm-ups; share = lane_share_pct (0 = library default: 50 % below\n"
      "32 frames, 25 % from 32 -- the setting tuned for overlapping forwards).")
for prec, share, batches in rows:
    m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision=prec)
    m.load_state_dict(sd)
    if share > 0:
        m.set_option("lane_share_pct", share)
    for B in batches:
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
        print("%-5s share %3s batch %3d: median %.3f ms (min %.3f) -> %.0f frames/s one call at a time" % (
            prec, "dflt" if share == 0 else "%d%%" % share, B, ts[10], ts[0], B / ts[10] * 1e3))
    del m
    torch.cuda.empty_cache()
