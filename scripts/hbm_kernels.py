#!/usr/bin/env python
"""Run the two HBM-bound kernels alone on inputs larger than L2 (for ncu captures): hbm_kernels.py [softargmax|fk]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import consts, synth
from hrp_b200.model import FkRobot, soft_argmax
which = sys.argv[1] if len(sys.argv) > 1 else "softargmax"
dev = torch.device("cuda", 0)
if which == "softargmax":
    B, nk = 64, 7
    hm = torch.randn(B, nk * 64, 64, 64, device=dev)
    K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_camera(B, 3))
    rz = torch.ones(B, device=dev)
    for _ in range(4):
        uvd, xyz = soft_argmax(hm, nk, K, rz, 1.3, 256.0, 3, True)
else:
    fk = FkRobot("panda")
    q, rot, tr, K = (torch.from_numpy(a).to(dev).repeat(100, *([1] * (a.ndim - 1))) for a in synth.make_fk_inputs("panda", 100_000, 1))
    for _ in range(4):
        xyz, uv = fk.keypoints(q, rot, tr, K)
torch.cuda.synchronize()
print("ok", which)
