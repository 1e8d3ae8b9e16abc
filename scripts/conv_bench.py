#!/usr/bin/env python
"""Time single tensor-core conv layers in isolation (hrp_conv_bench). Each line: shape, microseconds, TFLOP/s.
usage: conv_bench.py <bf16|tf32> [B] [shape ...]   shape = H,Cin,Cout,k,stride,res   (default: the HRNet/ResNet hot layers)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hrp_b200  # noqa
from hrp_b200 import capi
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
shapes = [tuple(int(v) for v in s.split(",")) for s in sys.argv[3:]] or [
    (64, 32, 32, 3, 1, 1), (64, 32, 32, 3, 1, 0), (32, 64, 64, 3, 1, 1), (16, 128, 128, 3, 1, 1), (8, 256, 256, 3, 1, 1),
    (64, 64, 64, 3, 1, 0), (64, 64, 256, 1, 1, 1), (64, 256, 64, 1, 1, 0), (32, 128, 512, 1, 1, 1), (16, 256, 1024, 1, 1, 1)]
torch.cuda.init()
torch.zeros(1, device="cuda")
ms = C.c_float()
for H, Cin, Cout, k, stride, res in shapes:
    capi.check(capi.lib().hrp_conv_bench(capi.PREC[prec], B, H, H, Cin, Cout, k, stride, res, 20, C.byref(ms), None))
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    fl = 2.0 * B * Ho * Ho * Cout * k * k * Cin * (8 if res == 8 else 1)      # res == 8: the 8-conv branch chain kernel
    print("%s B=%d %3dx%-3d %4d->%-4d k%d s%d res%d  %8.2f us  %7.1f TF/s" % (prec, B, H, H, Cin, Cout, k, stride, res, ms.value * 1e3, fl / ms.value / 1e9))
