#!/usr/bin/env python
"""Development probe for the tensor-core conv family: runs a ladder of shapes through hrp_conv2d_nhwc and prints where
(rows / channels / k-blocks) the result departs from a float64 reference. Not a test; run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hrp_b200  # noqa: F401,E402
from hrp_b200.model import conv2d_nhwc  # noqa: E402


def rnd(x, prec):
    if prec == "bf16":
        return x.bfloat16().float()
    if prec == "tf32":
        return ((x.view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)
    return x


def run(prec, B, H, Cin, Cout, k, stride, res=True, relu=True, seed=0):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    x = rnd(torch.randn(B, H, H, Cin, generator=g), prec)
    w = rnd(torch.randn(Cout, Cin, k, k, generator=g) / (k * k * Cin) ** 0.5, prec)
    b = torch.randn(Cout, generator=g)
    pad = k // 2
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, pad)
    r = rnd(torch.randn(ref.shape, generator=g), prec) if res else None
    if res:
        ref = ref + r.double()
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), r.permute(0, 2, 3, 1).contiguous().to(dev) if res else None,
                      stride, pad, relu, prec).cpu().double()
    err = (out - ref).abs()
    tol = {"bf16": 2.0 ** -7, "tf32": 2.0 ** -10, "fp32": 1e-5}[prec] * (1.0 + ref.abs())
    bad = err > tol
    print("%-5s B=%d H=%d Cin=%d Cout=%d k=%d s=%d res=%d: max err %.3e  bad %.4f%%" % (
        prec, B, H, Cin, Cout, k, stride, res, float(err.max()), 100.0 * float(bad.float().mean())), flush=True)
    if bad.any():
        e2 = bad.reshape(-1, Cout)
        rows = e2.any(1).nonzero().flatten()
        cols = e2.any(0).nonzero().flatten()
        print("   bad rows %d/%d first %s ; bad cols %d/%d first %s" % (len(rows), e2.shape[0], rows[:12].tolist(), len(cols), Cout, cols[:12].tolist()))
        print("   out[0,0,0,:8]", out[0, 0, 0, :8].tolist())
        print("   ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
    return not bool(bad.any())


if __name__ == "__main__":
    precs = sys.argv[1:] or ["bf16", "tf32"]
    ok = True
    for prec in precs:
        ladder = [
            (1, 16, 64, 32, 1, 1, False, False),     # M=256, one k-block (bf16) / two (tf32)
            (1, 16, 64, 32, 1, 1, True, True),
            (1, 16, 256, 64, 1, 1, True, True),      # 4 k-blocks
            (2, 8, 512, 256, 1, 1, True, True),      # wide N
            (2, 16, 32, 32, 3, 1, True, True),       # 3x3, Cin=32 (two taps per k-block in bf16; partial last block)
            (2, 64, 32, 32, 3, 1, True, True),
            (3, 32, 64, 64, 3, 1, True, True),
            (2, 16, 128, 128, 3, 1, True, True),
            (2, 8, 256, 256, 3, 1, True, True),
            (2, 32, 32, 64, 3, 2, True, True),
            (1, 8, 1024, 2048, 1, 1, True, True),
            (5, 17, 64, 96, 3, 2, True, True),       # ragged M, Cout = 32*3
            (2, 64, 256, 448, 1, 1, False, False),   # heatmap conv shape
            (64, 8, 512, 2048, 1, 1, True, True),    # BLOCK_N = 256
            (64, 8, 512, 512, 3, 1, True, True),     # K = 4608: many barrier phase wraps
            (16, 16, 1024, 512, 1, 2, True, True),   # 1x1 stride 2 (ResNet downsample)
            (64, 16, 256, 256, 3, 2, True, True),
            (64, 32, 64, 128, 3, 1, True, True),     # BLOCK_N = 128
        ]
        for c in ladder:
            try:
                ok &= run(prec, *c[:6], res=c[6], relu=c[7])
            except Exception as e:  # noqa: BLE001
                ok = False
                print("   EXCEPTION", prec, c, e, flush=True)
                break
    print("ALL OK" if ok else "FAILURES")
