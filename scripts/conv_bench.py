#!/usr/bin/env python
"""One conv layer through hrp_conv2d_nhwc, a few times (for ncu captures of conv_tc_kernel / conv_igemm_f32_kernel).
usage: conv_bench.py <prec> B H Cin Cout k stride [res]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hrp_b200  # noqa
from hrp_b200.model import conv2d_nhwc
prec = sys.argv[1]
B, H, Cin, Cout, k, stride = (int(v) for v in sys.argv[2:8])
res = len(sys.argv) > 8 and sys.argv[8] == "1"
dev = torch.device("cuda", 0)
x = torch.randn(B, H, H, Cin, device=dev)
w = torch.randn(Cout, Cin, k, k, device=dev) / (k * k * Cin) ** 0.5
b = torch.randn(Cout, device=dev)
Ho = (H + 2 * (k // 2) - k) // stride + 1
r = torch.randn(B, Ho, Ho, Cout, device=dev) if res else None
for _ in range(3):
    y = conv2d_nhwc(x, w, b, r, stride, k // 2, True, prec)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().mean()))
