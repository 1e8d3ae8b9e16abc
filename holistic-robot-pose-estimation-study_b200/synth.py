"""Deterministic synthetic weights and inputs (no network access: no checkpoints, no datasets).

Weights are a reference-format `state_dict` (names/shapes from arch.py) drawn per tensor from numpy PCG64 streams
seeded by (seed, crc32(name)), so every machine with numpy produces identical bytes. A plain He init in eval mode
explodes through ~100 residual/fusion additions, so BatchNorm running statistics come from a one-off
calibration pass over seeded noise images (scripts/make_bn_calib.py, stored in data/bn_calib_v2.npz); BN affine
parameters, conv biases and head scales are drawn so that every stage of the forward is well-scaled and every fused
epilogue term (BN fold, conv bias, residual, head bias) is exercised with non-trivial values. The last BatchNorm of
every residual block gets a small gain (U(0.05, 0.2), the usual "zero-init residual" practice in softened form): a
random ReLU network with unit-gain residual branches is chaotic -- it amplifies a 2^-9 operand rounding to ~10 % of
the feature norm over ~55 layers, which no trained checkpoint does -- and tolerances stated on such weights would say
nothing about the arithmetic. With damped branches the network is as well conditioned as SURVEY.md Appendix B assumes.

Inputs follow SURVEY.md §8d: images ~ U[0,1) fp32 NCHW (the /255 convention of scripts/test.py:93), pinhole K with
fx=fy in U(300,650), principal point 128 +- 8, and k_value = sqrt(fx*fy*1e6/area) (lib/core/function.py:107-110).
"""
import math
import os
import re
import zlib

import numpy as np

from . import arch, consts

CALIB_FILE = os.path.join(consts.DATA_DIR, "bn_calib_v2.npz")
CALIB_FILE_UNDAMPED = os.path.join(consts.DATA_DIR, "bn_calib_undamped.npz")   # recipe="undamped" (SURVEY 8d as written)
_RESIDUAL_TAIL = re.compile(r"(layer\d\.\d+\.bn3|incre_modules\.\d\.0\.bn3|branches\.\d\.\d\.bn2)\.weight$")
CALIB_IMAGE_SEED = 7
CALIB_BATCH = 8


def _rng(seed, name):
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(name.encode())]))


def _draw(name, shape, kind, seed, robot, damped=True):
    g = _rng(seed, name)
    f32 = np.float32
    if kind == "conv_w":
        cout, cin, kh, kw = shape
        std = math.sqrt(2.0 / (kh * kw * cout))          # ctor init, full_net.py:178-181
        if name == "depth_layer.weight":
            std = 1e-3                                      # full_net.py:185-188
        elif name.endswith("final_layer.weight"):
            std = 0.1 if cin == 256 else 0.05               # heatmap logits sigma ~ 1: a real peak, far from one-hot
        return (g.standard_normal(shape) * std).astype(f32)
    if kind == "deconv_w":
        cin, cout, kh, kw = shape
        return (g.standard_normal(shape) * math.sqrt(2.0 / (cin * 4))).astype(f32)  # 2x2 taps reach each output
    if kind == "conv_b":
        if name == "depth_layer.bias":
            return np.full(shape, 0.8, f32)                 # depth ~ 0.8*k/1000 m
        return (g.standard_normal(shape) * 0.05).astype(f32)
    if kind == "bn_w":
        if damped and _RESIDUAL_TAIL.search(name):
            return g.uniform(0.05, 0.2, shape).astype(f32)  # damped residual branch (module docstring)
        return g.uniform(0.6, 1.4, shape).astype(f32)
    if kind == "bn_b":
        return (g.standard_normal(shape) * 0.2).astype(f32)
    if kind == "bn_mean":
        return np.zeros(shape, f32)
    if kind == "bn_var":
        return np.ones(shape, f32)
    if kind == "bn_nbt":
        return np.asarray(1, np.int64)
    if kind == "lin_w":
        cout, cin = shape
        if name.startswith("dec"):
            return (g.standard_normal(shape) * 0.02).astype(f32)   # ~0.1 rad per refinement iteration
        b = 1.0 / math.sqrt(cin)
        return g.uniform(-b, b, shape).astype(f32)
    if kind == "lin_b":
        return (g.standard_normal(shape) * 0.02).astype(f32)
    if kind == "buf":
        if name == "init_pose":
            return np.asarray([consts.ROBOTS[robot]["init_pose"]], f32)
        return np.asarray([consts.INIT_ROT6D], f32)
    raise ValueError(kind)


_calib_cache = {}


def _calib(recipe):
    if recipe not in _calib_cache:
        path = CALIB_FILE if recipe == "damped" else CALIB_FILE_UNDAMPED
        _calib_cache[recipe] = np.load(path) if os.path.exists(path) else None
    return _calib_cache[recipe]


def make_state_dict(robot, backbone="resnet50", seed=1234, calibrated=True, recipe="damped", ctor=None):
    """Ordered dict name -> numpy array in the reference's state-dict format.

    recipe "damped" (default, module docstring) or "undamped": every BatchNorm gain U(0.6, 1.4), i.e. SURVEY.md 8d's
    recipe as written. The undamped network is chaotic (it amplifies operand rounding); it is kept to REPORT how far each
    precision family is from the gates there and to exercise BN folding on different running statistics."""
    if recipe not in ("damped", "undamped"):
        raise ValueError(recipe)
    variant = "resnet50" if backbone in ("resnet", "resnet50") else "hrnet32"
    z = _calib(recipe) if calibrated else None
    if calibrated and z is None:
        raise FileNotFoundError("BN calibration file for recipe %r missing: run scripts/make_bn_calib.py" % recipe)
    sd = {}
    for name, shape, kind in arch.full_net(robot, backbone, ctor):
        t = _draw(name, shape, kind, seed, robot, damped=recipe == "damped")
        if name.startswith(("depth_bn.", "joint_conv_layers.")) and kind in ("bn_mean", "bn_var"):   # variant-only BatchNorms: drawn, not calibrated
            g = _rng(seed, name)
            t = (g.standard_normal(shape) * 0.1 if kind == "bn_mean" else g.uniform(0.5, 1.5, shape)).astype(np.float32)
        elif z is not None and kind in ("bn_mean", "bn_var"):
            key = "%d/%s/%s" % (seed, variant if name.startswith(("reg_backbone", "deconv")) else "rootnet", name)
            t = z[key].astype(np.float32)
            assert t.shape == tuple(shape), (name, t.shape, shape)
        sd[name] = t
    return sd


def make_pretrained_rootnet_state(sd):
    """A DepthNet pre-training checkpoint (lib/models/depth_net.py RootNet naming: `backbone.*`, `depth_layer.*`) that
    differs measurably from the rootnet tensors of `sd`: same convolutions, other depth head, last BN gain scaled."""
    out = {}
    for k, v in sd.items():
        if k.startswith("rootnet_backbone."):
            out["backbone." + k[len("rootnet_backbone."):]] = np.array(v, copy=True)
    out["backbone.final_feat_layer.1.weight"] = out["backbone.final_feat_layer.1.weight"] * np.float32(0.9)
    g = np.random.Generator(np.random.PCG64(77))
    out["depth_layer.weight"] = (sd["depth_layer.weight"] + 2e-4 * g.standard_normal(sd["depth_layer.weight"].shape)).astype(np.float32)
    out["depth_layer.bias"] = np.full_like(sd["depth_layer.bias"], 0.7)
    out["xy_layer.weight"] = np.zeros((1, 256, 1, 1), np.float32)      # a tensor the full network does not have (strict=False)
    return out


def make_images(batch, seed):
    g = np.random.Generator(np.random.PCG64([int(seed), 1]))
    return g.random((batch, 3, consts.IMAGE_SIZE, consts.IMAGE_SIZE), dtype=np.float32)


def make_camera(batch, seed):
    """(K [B,3,3], k_value [B]) fp32."""
    g = np.random.Generator(np.random.PCG64([int(seed), 2]))
    f = g.uniform(300.0, 650.0, batch)
    cx = 128.0 + g.uniform(-8.0, 8.0, batch)
    cy = 128.0 + g.uniform(-8.0, 8.0, batch)
    K = np.zeros((batch, 3, 3), np.float32)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = cx
    K[:, 1, 2] = cy
    K[:, 2, 2] = 1.0
    side = np.maximum(g.uniform(120.0, 256.0, batch), g.uniform(120.0, 256.0, batch))
    k_value = np.sqrt(f * f * 1000.0 * 1000.0 / (side * side)).astype(np.float32)
    return K, k_value


def make_inputs(batch, seed):
    """(images [B,3,256,256], K [B,3,3], k_value [B]) fp32 numpy."""
    K, kv = make_camera(batch, seed)
    return make_images(batch, seed), K, kv


def make_frames(n, seed, h=480, w=640):
    """Raw camera frames + boxes for the input-side op (hrp_crop_resize_u8): uint8 HWC [n,h,w,3] (smooth structure plus
    pixel noise, so bilinear weights matter), integer crop boxes inside the frame (smaller and larger than the 256-pixel
    crop, non-square), strict robot boxes inside them, and cameras in the range of the datasets' intrinsics."""
    g = np.random.Generator(np.random.PCG64([int(seed), 9]))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    frames = np.empty((n, h, w, 3), np.uint8)
    crop = np.empty((n, 4), np.int32)
    kbox = np.empty((n, 4), np.float32)
    K = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        img = np.zeros((h, w, 3), np.float32)
        for c in range(3):
            a, b2, ph = g.uniform(0.01, 0.06, 2), g.uniform(0.01, 0.05, 2), g.uniform(0, 6.28, 2)
            img[..., c] = 110 + 60 * np.sin(a[0] * xx + b2[0] * yy + ph[0]) + 40 * np.cos(a[1] * xx - b2[1] * yy + ph[1])
        img += g.normal(0, 12, img.shape)
        frames[i] = np.clip(img, 0, 255).astype(np.uint8)
        bw, bh = int(g.integers(90, 470)), int(g.integers(90, 470))
        x0, y0 = int(g.integers(0, w - bw + 1)), int(g.integers(0, h - bh + 1))
        crop[i] = (x0, y0, x0 + bw, y0 + bh)
        mx, my = g.uniform(0.05, 0.2, 2) * (bw, bh)
        kbox[i] = (x0 + mx, y0 + my, x0 + bw - mx * g.uniform(0.5, 1.5), y0 + bh - my * g.uniform(0.5, 1.5))
        f = g.uniform(500.0, 650.0)
        K[i] = [[f, 0, w / 2 + g.uniform(-10, 10)], [0, f * g.uniform(0.98, 1.02), h / 2 + g.uniform(-10, 10)], [0, 0, 1]]
    crop[0] = (100, 50, 356, 306)                # exactly 256 x 256: the reference skips the resize (augmentations.py:193-195)
    kbox[0] = (120.0, 70.0, 330.0, 290.0)
    return frames, crop, kbox, K


def make_fk_inputs(robot, n, seed):
    """Pose-sweep inputs: q ~ U(JOINT_BOUNDS), noisy rot6d of a random rotation, trans, K."""
    spec = consts.ROBOTS[robot]
    g = np.random.Generator(np.random.PCG64([int(seed), 3]))
    b = np.asarray(spec["bounds"], np.float64)
    q = (b[:, 0] + (b[:, 1] - b[:, 0]) * g.random((n, spec["dof"]))).astype(np.float32)
    a = g.standard_normal((n, 3))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    c = g.standard_normal((n, 3))
    c -= (c * a).sum(1, keepdims=True) * a
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    rot6d = (np.concatenate([a, c], 1) + 1e-2 * g.standard_normal((n, 6))).astype(np.float32)
    trans = np.stack([g.uniform(-0.5, 0.5, n), g.uniform(-0.5, 0.5, n), g.uniform(0.6, 2.0, n)], 1).astype(np.float32)
    K, _ = make_camera(n, seed + 17)
    return q, rot6d, trans, K


def make_metrics_inputs(robot, n, seed):
    """Evaluation-tail inputs: a ground-truth pose sweep and a prediction = ground truth + noise (a few
    centimetres / hundredths of a radian, growing along the batch so the ADD curve has a shape), the ORIGINAL camera (640x480).
    Returns dict of float32 arrays: gt_q, gt_rot, gt_trans, q, rot, trans, K."""
    gq, grot, gtr, _ = make_fk_inputs(robot, n, seed)
    g = np.random.Generator(np.random.PCG64([int(seed), 11]))
    amp = np.linspace(0.05, 1.5, n)[:, None]
    q = (gq + 0.04 * amp * g.standard_normal(gq.shape)).astype(np.float32)
    rot = (grot + 0.02 * amp * g.standard_normal(grot.shape)).astype(np.float32)
    trans = (gtr + 0.03 * amp * g.standard_normal(gtr.shape)).astype(np.float32)
    K = np.zeros((n, 3, 3), np.float32)
    K[:, 0, 0] = g.uniform(500.0, 650.0, n); K[:, 1, 1] = K[:, 0, 0] * g.uniform(0.98, 1.02, n)
    K[:, 0, 2] = g.uniform(300.0, 340.0, n); K[:, 1, 2] = g.uniform(220.0, 260.0, n); K[:, 2, 2] = 1.0
    return dict(gt_q=gq, gt_rot=grot, gt_trans=gtr, q=q, rot=rot, trans=trans, K=K)


def make_heatmaps(batch, nkpt, seed, mode="blobs"):
    """Adversarial heatmap logits [B, nkpt*64, 64, 64] fp32 for kernel-level soft-argmax tests."""
    g = np.random.Generator(np.random.PCG64([int(seed), 4]))
    D = consts.DEPTH_DIM
    x = g.standard_normal((batch, nkpt, D, D, D), dtype=np.float32)
    if mode == "noise":
        pass
    elif mode == "blobs":       # Gaussian blob x20 + noise, peaks allowed at the borders
        ax = np.arange(D, dtype=np.float32)
        for b in range(batch):
            for k in range(nkpt):
                c = g.uniform(-2.0, D + 1.0, 3)
                s = g.uniform(1.0, 4.0)
                gz = np.exp(-0.5 * ((ax - c[0]) / s) ** 2)
                gy = np.exp(-0.5 * ((ax - c[1]) / s) ** 2)
                gx = np.exp(-0.5 * ((ax - c[2]) / s) ** 2)
                x[b, k] += 20.0 * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
    elif mode == "extreme":     # large-magnitude logits: exercises the max subtraction
        x *= 30.0
        x += g.uniform(-80.0, 80.0, (batch, nkpt, 1, 1, 1)).astype(np.float32)
    else:
        raise ValueError(mode)
    return x.reshape(batch, nkpt * D, D, D)
