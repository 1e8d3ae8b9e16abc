#!/usr/bin/env python
"""Run the fused BasicBlock kernel a few times (for ncu captures of conv_block_kernel): block_bench.py [B] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200.model import basic_block_nhwc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, H, 32, generator=g).to(dev)
w1 = (torch.randn(32, 32, 3, 3, generator=g) / 17).to(dev)
w2 = (torch.randn(32, 32, 3, 3, generator=g) / 17).to(dev)
b = torch.randn(32, generator=g).to(dev)
for _ in range(4):
    y = basic_block_nhwc(x, w1, b, w2, b)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
