"""GPU: the CUDA path (through the C ABI) against the oracle, the committed reference goldens and size-independent
properties at BASELINE.json's full sizes. Tolerances are the north_star gates (helpers.TOL_*) for the end-to-end net and
tight fp32 bounds for the kernels (SURVEY.md §8d "Parity gates")."""
import numpy as np
import pytest
import torch

from hrp_b200 import consts, synth
from oracle import integral, kinematics

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---------------------------------------------------------------------------------------------------- FK + projection
@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_fk_against_reference_golden_and_oracle(robot, dev):
    from hrp_b200.model import FkRobot
    g = helpers.load_golden("fk_%s.npz" % robot)
    seed, n, root = (int(v) for v in g["meta"])
    q, rot, tr, K = synth.make_fk_inputs(robot, n, seed)
    fk = FkRobot(robot)
    xyz, uv = fk.keypoints(cu(q, dev), cu(rot, dev), cu(tr, dev), cu(K, dev))
    assert helpers.maxdiff(xyz, g["xyz"]) < 1e-5                     # metres
    ok = np.abs(g["xyz"][..., 2]) > 0.3          # d(uv) ~ f * d(xyz) / z^2: keypoints near the camera plane amplify fp32 noise
    assert np.max(np.abs(uv.cpu().numpy() - g["uv"])[ok]) < 1e-2     # pixels
    # ragged size (not a multiple of the CTA) and a single pose
    for m in (1, 129, 511):
        x2, u2 = fk.keypoints(cu(q[:m], dev), cu(rot[:m], dev), cu(tr[:m], dev), cu(K[:m], dev))
        assert torch.equal(x2, xyz[:m]) and torch.equal(u2, uv[:m])
    x0, _ = fk.keypoints(cu(q[:0], dev), cu(rot[:0], dev), cu(tr[:0], dev), cu(K[:0], dev))
    assert x0.shape == (0, fk.program.nkpt, 3)


@pytest.mark.parametrize("robot", ["panda", "kuka"])
def test_fk_limb_lengths_at_scale(robot, dev):
    """Size-independent property at sweep size: consecutive keypoint distances are the in-tree limb lengths
    (lib/dataset/const.py:108-124) for every pose, and keypoint[root] == trans (urdf_robot.py:218-222)."""
    from hrp_b200.model import FkRobot
    n = 1_000_000
    q, rot, tr, K = synth.make_fk_inputs(robot, n, 321)
    fk = FkRobot(robot)
    xyz, uv = fk.keypoints(cu(q, dev), cu(rot, dev), cu(tr, dev), cu(K, dev))
    d = (xyz[:, 1:] - xyz[:, :-1]).norm(dim=2)
    ref = torch.tensor(consts.LIMB_LENGTH[robot], device=dev)
    assert float((d - ref).abs().max()) < 5e-5
    root = consts.ROBOTS[robot]["ref_kp"]
    assert float((xyz[:, root] - cu(tr, dev)).abs().max()) < 1e-5
    assert torch.isfinite(uv).all()



@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_fk_generated_chain_matches_table_interpreter(robot, dev, monkeypatch):
    """The URDF-generated straight-line chain (persistent cp.async kernel) against the table interpreter on the same poses:
    many tiles per CTA, a ragged tail, and input views that are not 16-byte aligned (synchronous staging path)."""
    from hrp_b200.model import FkRobot
    n = 148 * 7 * 128 * 2 + 77
    args = [cu(a, dev) for a in synth.make_fk_inputs(robot, n, 55)]
    fk = FkRobot(robot)
    monkeypatch.setenv("HRP_FK_GENERIC", "1")
    fk_tab = FkRobot(robot)
    monkeypatch.delenv("HRP_FK_GENERIC")
    x1, u1 = fk.keypoints(*args)
    x0, u0 = fk_tab.keypoints(*args)
    assert helpers.maxdiff(x1, x0) < 2e-6
    ok = x0[..., 2].abs() > 0.3
    assert float((u1 - u0).abs()[ok].max()) < 1e-2
    shifted = [a[1:] for a in args]              # rot6d rows are 24 B, trans rows 12 B: misaligned for 16-byte copies
    x2, u2 = fk.keypoints(*shifted)
    assert torch.equal(x2, x1[1:]) and torch.equal(u2, u1[1:])

def test_fk_branching_tree_against_oracle(dev):
    """A URDF whose keypoint paths branch mid-chain (saved frames in shared memory), prismatic + off-axis joints."""
    import ctypes as C
    from hrp_b200 import capi, urdf
    from test_host import BRANCHY
    R = urdf.Robot(BRANCHY)
    frames = [("b", np.zeros(3)), ("l3", np.array([0.0, 0.0, 0.01])), ("r2", np.zeros(3)), ("tip", np.zeros(3))]
    orc = kinematics.OracleRobot(BRANCHY)
    rng = np.random.default_rng(3)
    n = 300
    q = rng.uniform(-2, 2, (n, 5)).astype(np.float32)
    _, rot, tr, K = synth.make_fk_inputs("panda", n, 8)
    for root in (0, 2):
        P = urdf.compile_program(R, frames, root)
        s, keep = capi.fk_program_struct(P)
        h = C.c_void_p()
        capi.check(capi.lib().hrp_fk_create(C.byref(s), C.byref(h)))
        xyz = torch.empty(n, 4, 3, device=dev)
        uv = torch.empty(n, 4, 2, device=dev)
        args = [cu(a, dev) for a in (q, rot, tr, K)]
        capi.check(capi.lib().hrp_fk_project(h, *[C.c_void_p(a.data_ptr()) for a in args], n, C.c_void_p(xyz.data_ptr()),
                                             C.c_void_p(uv.data_ptr()), C.c_void_p(0)))
        torch.cuda.synchronize()
        ref = kinematics.keypoints(orc, frames, q, rot, tr, root)
        assert helpers.maxdiff(xyz, ref) < 1e-5
        capi.lib().hrp_fk_destroy(h)


# ---------------------------------------------------------------------------------------------------- soft-argmax
@pytest.mark.parametrize("case", ["resnet50_k7_blobs", "resnet50_k7_extreme", "resnet50_k17_noise",
                                  "hrnet32_k8_blobs", "hrnet32_k17_extreme"])
def test_softargmax_against_reference_golden(case, dev):
    from hrp_b200.model import soft_argmax
    g = helpers.load_golden("softargmax_%s.npz" % case)
    seed, B, nkpt = (int(v) for v in g["meta"])
    path, mode = case.split("_")[0], case.split("_")[-1]
    hm = synth.make_heatmaps(B, nkpt, seed, mode)
    K, _ = synth.make_camera(B, seed)
    uvd, xyz = soft_argmax(cu(hm, dev), nkpt, cu(K, dev), cu(g["root_z"], dev), 1.3, 256.0, 3, True)
    # the reference's hrnet path skips the re-normalisation and carries ~1e-4 of its own fp32 softmax rounding
    tol = 1e-5 if path == "resnet50" else 2e-4
    assert helpers.maxdiff(uvd, g["uvd"]) < tol
    assert helpers.maxdiff(xyz, g["xyz"]) < 4 * tol
    # float64 ground truth: the single-pass kernel is at least as accurate as the reference
    p = torch.softmax(torch.from_numpy(hm).double().reshape(B, nkpt, -1), 2).reshape(B, nkpt, 64, 64, 64)
    r = torch.arange(64, dtype=torch.float64)
    exact = torch.stack([(p.sum((2, 3)) * r).sum(2), (p.sum((2, 4)) * r).sum(2), (p.sum((3, 4)) * r).sum(2)], 2) / 64 - 0.5
    exact[:, 3, 2] = 0
    assert helpers.maxdiff(uvd, exact) < 5e-6


def test_softargmax_properties_full_size(dev):
    """B=64 Panda-size heatmaps (470 MB): shift invariance, one-hot recovery, uvd-only mode."""
    from hrp_b200.model import soft_argmax
    B, nkpt = 64, 7
    g = torch.Generator(device="cpu").manual_seed(5)
    hm = torch.randn(B, nkpt * 64, 64, 64, generator=g).to(dev) * 3
    uvd, none = soft_argmax(hm, nkpt, rootid=3, fixroot=False)
    assert none is None and uvd.shape == (B, nkpt, 3)
    uvd2, _ = soft_argmax(hm + 37.5, nkpt, rootid=3, fixroot=False)
    assert float((uvd - uvd2).abs().max()) < 2e-5
    hot = torch.full((2, nkpt * 64, 64, 64), -50.0, device=dev)
    coords = [(0, 0, 0), (63, 63, 63), (5, 60, 17), (31, 32, 33), (1, 2, 3), (62, 0, 63), (10, 10, 10)]
    for k, (d, h, w) in enumerate(coords):
        hot[:, k * 64 + d, h, w] = 60.0
    u, _ = soft_argmax(hot, nkpt, fixroot=False)
    exp = torch.tensor([[w / 64 - 0.5, h / 64 - 0.5, d / 64 - 0.5] for d, h, w in coords], device=dev)
    assert float((u - exp).abs().max()) < 1e-6


# ---------------------------------------------------------------------------------------------------- single conv layers
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride", [(2, 64, 32, 32, 3, 1), (3, 32, 64, 64, 3, 1), (2, 16, 128, 128, 3, 1),
                                                    (2, 8, 256, 256, 3, 1), (2, 64, 64, 256, 1, 1), (2, 32, 32, 64, 3, 2),
                                                    (1, 8, 1024, 2048, 1, 1), (5, 17, 48, 80, 3, 2)])
def test_conv_layer_against_torch_fp32(B, H, Cin, Cout, k, stride, dev):
    from hrp_b200.model import conv2d_nhwc
    g = torch.Generator().manual_seed(B * 1000 + Cin)
    x = torch.randn(B, H, H, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (k * k * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    pad = k // 2
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, b, stride, pad)
    res = torch.randn(ref.shape, generator=g)
    ref = torch.relu(ref + res).permute(0, 2, 3, 1).contiguous()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), res.permute(0, 2, 3, 1).contiguous().to(dev), stride, pad, True)
    assert helpers.maxdiff(out, ref) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,Cout,relu", [(64, 2048, 2048, False), (1, 1024, 1024, False), (5, 1024, 1024, True), (130, 2048, 1024, False),
                                             (64, 192, 16, True)])
def test_linear_few_rows_against_torch_fp32(B, Cin, Cout, relu, dev):
    """nn.Linear on the batch rows (the fc stacks of the regression heads, full_net.py:214-238): the skinny fp32 kernel
    against torch, and every row independent of how many rows travel with it."""
    from hrp_b200.model import conv2d_nhwc
    g = torch.Generator().manual_seed(B + Cin)
    x = torch.randn(B, 1, 1, Cin, generator=g)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.linear(x.view(B, Cin).double(), w.view(Cout, Cin).double(), b.double())
    ref = (torch.relu(ref) if relu else ref).float()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), None, 1, 0, relu).view(B, Cout)
    assert helpers.maxdiff(out, ref) < 2e-5 * max(1.0, float(ref.abs().max()))
    one = conv2d_nhwc(x[B // 2:B // 2 + 1].to(dev), w.to(dev), b.to(dev), None, 1, 0, relu).view(1, Cout)
    assert torch.equal(one, out[B // 2:B // 2 + 1])


def _round_like(x, prec):
    if prec == "bf16":
        return x.bfloat16().float()
    if prec == "f16":
        return x.half().float()
    return ((x.view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)       # cvt.rna.tf32.f32


@pytest.mark.parametrize("prec", ["bf16", "tf32", "f16"])
@pytest.mark.parametrize("B,H,Cin,Cout,k,stride", [(1, 16, 64, 32, 1, 1), (2, 64, 32, 32, 3, 1), (3, 32, 64, 64, 3, 1),
                                                    (2, 16, 128, 128, 3, 1), (2, 8, 256, 256, 3, 1), (2, 32, 32, 64, 3, 2),
                                                    (1, 8, 1024, 2048, 1, 1), (5, 17, 64, 96, 3, 2), (2, 64, 256, 448, 1, 1),
                                                    (40, 64, 64, 256, 1, 1), (9, 32, 128, 512, 1, 1)])
def test_conv_layer_tensor_core_families(prec, B, H, Cin, Cout, k, stride, dev):
    """tcgen05 implicit GEMM (conv_tc.cu) against a float64 conv on operands rounded to the family's type: the only
    differences left are fp32 accumulation order and the rounding of the stored output (bf16: 2^-8, TF32 / IEEE half: 2^-11
    rel)."""
    from hrp_b200.model import conv2d_nhwc
    g = torch.Generator().manual_seed(B * 1000 + Cin + k)
    x = _round_like(torch.randn(B, H, H, Cin, generator=g), prec)
    w = _round_like(torch.randn(Cout, Cin, k, k, generator=g) / (k * k * Cin) ** 0.5, prec)
    b = torch.randn(Cout, generator=g)
    pad = k // 2
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, pad)
    res = _round_like(torch.randn(ref.shape, generator=g), prec)
    ref = torch.relu(ref + res.double()).permute(0, 2, 3, 1).contiguous()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), res.permute(0, 2, 3, 1).contiguous().to(dev), stride, pad, True, prec)
    tol = (2.0 ** -8 if prec == "bf16" else 2.0 ** -11) * (1.0 + ref.abs()) * 1.01 + 2e-5      # tf32 / f16: 11-bit significand
    assert bool(((out.cpu().double() - ref).abs() <= tol).all())



@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,res", [(2, 64, 32, 32, 3, 1, True), (3, 32, 64, 64, 3, 1, True), (2, 16, 128, 128, 3, 1, False),
                                                        (2, 8, 256, 256, 3, 1, True), (2, 32, 32, 64, 3, 2, True), (1, 8, 1024, 2048, 1, 1, False),
                                                        (5, 17, 64, 96, 3, 2, True), (2, 64, 32, 448, 1, 1, False), (9, 32, 128, 512, 1, 1, True),
                                                        (3, 64, 256, 64, 1, 1, False)])
def test_conv_layer_3xtf32_is_fp32_grade(B, H, Cin, Cout, k, stride, res, dev):
    """3xTF32 (conv_tc.cu: operands split into hi + lo TF32 halves inside the kernel, three tcgen05 products per k-step)
    on UNROUNDED fp32 operands against a float64 conv: fp32-grade, i.e. ~2^-20 relative instead of TF32's 2^-11 -- both
    operand paths (TMA boxes + splitter warps, cp.async gather), strided and 1x1 layers, with and without residual."""
    from hrp_b200.model import conv2d_nhwc
    g = torch.Generator().manual_seed(B * 1000 + Cin + k + 3)
    x = torch.randn(B, H, H, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (k * k * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    pad = k // 2
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride, pad)
    r = torch.randn(ref.shape, generator=g) if res else None
    if res:
        ref = ref + r.double()
    ref = torch.relu(ref).permute(0, 2, 3, 1).contiguous()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), r.permute(0, 2, 3, 1).contiguous().to(dev) if res else None, stride, pad, True, "tf32x3")
    err = (out.cpu().double() - ref).abs()
    assert float(err.max()) < 2e-5 * max(1.0, float(ref.abs().max())), float(err.max())      # the fp32 family's own bound
    single = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), r.permute(0, 2, 3, 1).contiguous().to(dev) if res else None, stride, pad, True, "tf32")
    assert float(err.max()) < 0.05 * float((single.cpu().double() - ref).abs().max())          # and far better than one TF32 pass


@pytest.mark.parametrize("prec", ["bf16", "tf32", "f16"])
@pytest.mark.parametrize("B,H,C,res,relu", [(37, 64, 32, True, True), (9, 32, 64, False, True), (3, 24, 32, True, False),
                                             (5, 40, 64, False, False), (1, 8, 32, True, True), (130, 16, 32, False, True),
                                             (20, 16, 128, True, True), (64, 8, 256, True, True), (3, 32, 128, False, True),
                                             (5, 16, 256, False, False), (1, 8, 128, True, False), (70, 16, 128, False, True)])
def test_conv3x3_shifted_gemm_kernel(prec, B, H, C, res, relu, dev):
    """3x3/s1/p1 convs with Cin == Cout: conv_slab.cu for 32 / 64 channels (units that straddle frames, several units per
    persistent CTA, map widths that are not a multiple of 8 positions), conv_tc.cu for the wider ones; with and without
    residual / ReLU. Same exactness bound as above."""
    from hrp_b200.model import conv2d_nhwc
    g = torch.Generator().manual_seed(B * 100 + H + C)
    x = _round_like(torch.randn(B, H, H, C, generator=g), prec)
    w = _round_like(torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5, prec)
    b = torch.randn(C, generator=g)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), 1, 1)
    r = _round_like(torch.randn(ref.shape, generator=g), prec) if res else None
    if res:
        ref = ref + r.double()
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).contiguous()
    out = conv2d_nhwc(x.to(dev), w.to(dev), b.to(dev), r.permute(0, 2, 3, 1).contiguous().to(dev) if res else None, 1, 1, relu, prec)
    tol = (2.0 ** -8 if prec == "bf16" else 2.0 ** -11) * (1.0 + ref.abs()) * 1.01 + 2e-5      # tf32 / f16: 11-bit significand
    assert bool(((out.cpu().double() - ref).abs() <= tol).all())


@pytest.mark.parametrize("B,H", [(37, 64), (3, 64), (64, 64), (1, 64), (9, 32), (130, 16)])
def test_fused_basic_block_equals_two_convs(B, H, dev):
    """conv_block.cu (a whole HRNet BasicBlock in one launch, intermediate in shared memory) against the same block run
    as two tensor-core convs: same operands, same K order, same bf16 rounding of the intermediate => bit-identical."""
    from hrp_b200.model import basic_block_nhwc, conv2d_nhwc
    C = 32
    g = torch.Generator().manual_seed(B * 7 + H)
    x = torch.randn(B, H, H, C, generator=g).bfloat16().float().to(dev)
    w1 = (torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev)
    w2 = (torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev)
    b1, b2 = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    y = conv2d_nhwc(x, w1, b1, None, 1, 1, True, "bf16")
    ref = conv2d_nhwc(y, w2, b2, x, 1, 1, True, "bf16")
    out = basic_block_nhwc(x, w1, b1, w2, b2)
    assert torch.equal(out, ref), float((out - ref).abs().max())


@pytest.mark.parametrize("B,H,C,nblocks", [(64, 16, 128, 4), (5, 16, 128, 1), (64, 8, 256, 4), (3, 8, 256, 2), (301, 8, 256, 1),
                                            (160, 16, 128, 2), (2, 4, 256, 3), (7, 12, 128, 4),
                                            (64, 32, 64, 4), (3, 32, 64, 1), (150, 32, 64, 2), (5, 32, 64, 3), (2, 16, 64, 4), (9, 24, 64, 4),
                                            (4, 5, 64, 2), (300, 8, 64, 1)])
def test_branch_chain_equals_layer_by_layer(B, H, C, nblocks, dev):
    """conv_chain.cu (all BasicBlocks of a low-resolution HRNet branch in one launch, one image per CTA, activations in
    shared memory; C = 128 / 256) and conv_roll.cu (C = 64: a single shared-memory buffer that every conv rewrites in place
    a few rows further up; several images per CTA, one / two / three passes per conv, maps that leave grid positions behind
    the last tile) against the same blocks run conv by conv: same operands, K order and bf16 roundings => bit-identical."""
    from hrp_b200.model import basic_chain_nhwc, conv2d_nhwc
    g = torch.Generator().manual_seed(B * 11 + H + C)
    x = torch.randn(B, H, H, C, generator=g).bfloat16().float().to(dev)
    w = (torch.randn(2 * nblocks, C, C, 3, 3, generator=g) / (9 * C) ** 0.5 * 0.8).to(dev)
    b = (0.3 * torch.randn(2 * nblocks, C, generator=g)).to(dev)
    ref = x
    for k in range(nblocks):
        y = conv2d_nhwc(ref, w[2 * k], b[2 * k], None, 1, 1, True, "bf16")
        ref = conv2d_nhwc(y, w[2 * k + 1], b[2 * k + 1], ref, 1, 1, True, "bf16")
    out = basic_chain_nhwc(x, w, b)
    assert torch.equal(out, ref), float((out - ref).abs().max())


# ---------------------------------------------------------------------------------------------------- full network
_models = {}


def gpu_model(robot, backbone, dev, precision="fp32"):
    from hrp_b200.model import HoliRobPoseB200
    key = (robot, backbone, precision)
    if key not in _models:
        _, sd = helpers.oracle_for(robot, backbone)
        m = HoliRobPoseB200(robot, {"backbone_name": backbone}, device=dev, precision=precision)
        m.load_state_dict(sd)
        _models[key] = m
    return _models[key]


def check_gates(out, ref, scale=1.0):
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk", "kp2d_int", "kp2d_fk"]
    d = {k: helpers.maxdiff(out[k], ref[k]) for k in names}
    assert d["joint_angles"] < helpers.TOL_RAD * scale, d
    assert d["root_depth"] < helpers.TOL_DEPTH_M * scale, d
    assert max(d["kp2d_int"], d["kp2d_fk"], d["root_uv"]) < helpers.TOL_PX * scale, d
    assert d["rot6d"] < 1e-3 * scale and d["uvd"] < 1e-3 * scale and max(d["kp3d_int"], d["kp3d_fk"], d["trans"]) < 2e-3 * scale, d
    return d


@pytest.mark.parametrize("robot,backbone", helpers.FULLNET_CASES)
def test_fullnet_against_reference_golden(robot, backbone, dev):
    g = helpers.load_golden("fullnet_%s_%s.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    m = gpu_model(robot, backbone, dev)
    img, K, kv = helpers.inputs(B, seed)
    tup = m(img.to(dev), img.to(dev), kv.to(dev), K.to(dev))
    assert len(tup) == 8 and all(t.dtype == torch.float32 and t.is_cuda for t in tup)
    assert [tuple(t.shape) for t in tup] == [g[k].shape for k in ("joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk")]
    out = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    d = check_gates(out, g)
    # fp32 family: far inside the gates
    assert d["joint_angles"] < 1e-4 and d["root_depth"] < 1e-4 and d["kp2d_fk"] < 0.05, d
    assert helpers.maxdiff(m.debug_tensor("xf", B).view(B, -1), g["probe_xf"]) < 1e-3
    assert helpers.maxdiff(m.debug_tensor("img_feat", B).view(B, -1), g["probe_img_feat"]) < 1e-3
    lg = m.debug_tensor("logits", B).view(B, -1, 64, 64)
    assert helpers.maxdiff(lg[:, ::37, ::5, ::7], g["probe_logits_sample"]) < 5e-3
    assert m.launch_count() > 300


def test_fullnet_graph_replay_and_eager_agree(dev):
    m = gpu_model("panda", "resnet50", dev)
    img, K, kv = helpers.inputs(3, 77)
    a = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    b = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))      # graph replay
    m.set_option("cuda_graph", 0)
    c = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))      # eager launches, caller pointers
    m.set_option("cuda_graph", 1)
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k



@pytest.mark.parametrize("backbone,prec", [("resnet50", "bf16"), ("hrnet32", "bf16"), ("hrnet32", "fp32"), ("resnet50", "f16")])
def test_fullnet_multi_lane_graph_equals_serial_execution(backbone, prec, dev):
    """The captured graph runs HRNet branches, the two backbones and the heads on separate streams (lane arenas + event
    edges); at a batch that really overlaps them it must reproduce the serial single-stream execution bit for bit."""
    m = gpu_model("panda", backbone, dev, prec)
    B = 24
    img, K, kv = helpers.inputs(B, 4141)
    img, K, kv = img.to(dev), K.to(dev), kv.to(dev)
    m.set_option("cuda_graph", 0)
    serial = {k: v.clone() for k, v in m.forward_dict(img, K, kv).items()}
    m.set_option("cuda_graph", 1)
    for _ in range(3):
        out = m.forward_dict(img, K, kv)
        for k in out:
            assert torch.equal(out[k], serial[k]), k

def test_fullnet_batch64_frames_are_independent(dev):
    """BASELINE config 2 size (Panda, B=64): every frame of the big batch equals the same frame run alone or in a small
    batch (eval-mode BN, per-frame softmax/FK => no cross-frame coupling), and the oracle agrees on a sample of frames."""
    m = gpu_model("panda", "resnet50", dev)
    img, K, kv = helpers.inputs(64, 909)
    big = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    idx = [0, 17, 63]
    small = m.forward_dict(img[idx].to(dev), K[idx].to(dev), kv[idx].to(dev))
    for k in big:
        assert helpers.maxdiff(big[k][idx], small[k]) < 1e-5, k
    om, _ = helpers.oracle_for("panda", "resnet50")
    ref = om.forward_dict(img[idx], img[idx], kv[idx], K[idx])
    check_gates({k: v[idx] for k, v in big.items()}, ref)
    # default k_value path of the dict convenience (scripts/real_test.py:285-289)
    auto = m(img[:2].to(dev), K[:2].to(dev))
    kv2 = torch.sqrt(K[:2, 0, 0] * K[:2, 1, 1] * 1e6 / 256.0 ** 2)
    ref2 = om.forward_dict(img[:2], img[:2], kv2, K[:2])
    check_gates(auto, ref2)




def test_fullnet_bf16_branch_chain_and_small_batch_form_agree(dev):
    """bf16 family: at batch 64 the low-resolution HRNet branches run through the one-launch chain kernel, at batch 3 the
    executor runs the same convs layer by layer (one CTA per image would leave the GPU idle). Both forms are bit-identical
    per layer, so a frame of the big batch must equal the same frame in the small batch."""
    m = gpu_model("panda", "resnet50", dev, "bf16")
    img, K, kv = helpers.inputs(64, 911)
    big = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    idx = [1, 30, 62]
    small = m.forward_dict(img[idx].to(dev), K[idx].to(dev), kv[idx].to(dev))
    for k in big:
        assert helpers.maxdiff(big[k][idx], small[k]) < 1e-5, k


def test_fullnet_empty_and_single_frame_batches(dev):
    """Edge sizes of the batch axis: an empty batch comes back as empty fields of the right trailing shapes (what the
    reference's torch modules do), one frame equals the same frame inside a batch of five, and shape errors are loud."""
    m = gpu_model("panda", "resnet50", dev)
    img, K, kv = helpers.inputs(5, 77)
    out0 = m.forward_dict(img[:0].to(dev), K[:0].to(dev), kv[:0].to(dev))
    assert out0["joint_angles"].shape == (0, m.dof) and out0["kp2d_fk"].shape == (0, m.nkpt, 2) and out0["kp3d_int"].shape == (0, m.nkpt, 3)
    t8 = m(img[:0].to(dev), img[:0].to(dev), kv[:0].to(dev), K=K[:0].to(dev))
    assert len(t8) == 8 and all(t.shape[0] == 0 for t in t8)
    five = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    one = m.forward_dict(img[3:4].to(dev), K[3:4].to(dev), kv[3:4].to(dev))
    for k in five:
        assert helpers.maxdiff(five[k][3:4], one[k]) < 1e-5, k
    with pytest.raises(ValueError):
        m.forward_dict(img[:, :, :128].to(dev), K.to(dev), kv.to(dev))
    with pytest.raises(ValueError):
        m.forward_dict(img, K.to(dev), kv.to(dev))          # host tensor: the caller moves data, never a silent copy


def test_fullnet_distinct_reg_and_root_images(dev):
    """The reference call takes two crops, x_reg (keypoint branch) and x_root (DepthNet), full_net.py:262-266; the callers
    usually pass the same tensor, which the graph path special-cases. Distinct tensors take the other graph and must
    match the oracle too -- and the 8-tuple interface returns the same values as the dict interface."""
    m = gpu_model("panda", "resnet50", dev)
    om, _ = helpers.oracle_for("panda", "resnet50")
    img, K, kv = helpers.inputs(2, 777)
    img2, _, _ = helpers.inputs(2, 778)
    tup = m(img.to(dev), img2.to(dev), kv.to(dev), K=K.to(dev))
    assert len(tup) == 8
    ref = om.forward_dict(img, img2, kv, K)
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk"]
    got = dict(zip(names, tup))
    assert helpers.maxdiff(got["joint_angles"], ref["joint_angles"]) < helpers.TOL_RAD
    assert helpers.maxdiff(got["root_depth"], ref["root_depth"]) < helpers.TOL_DEPTH_M
    assert helpers.maxdiff(got["root_uv"], ref["root_uv"]) < helpers.TOL_PX
    for k in ("trans", "kp3d_int", "kp3d_fk", "uvd", "rot6d"):
        assert helpers.maxdiff(got[k], ref[k]) < 1e-3, k
    # swapping the two crops changes the result (the two graphs really read different buffers)
    swapped = dict(zip(names, m(img2.to(dev), img.to(dev), kv.to(dev), K=K.to(dev))))
    assert helpers.maxdiff(swapped["root_depth"], got["root_depth"]) > 1e-4


@pytest.mark.parametrize("robot,backbone", [("panda", "resnet50"), ("baxter", "hrnet32")])
def test_fused_heatmap_head_matches_logits_plus_softargmax(robot, backbone, dev, monkeypatch):
    """Tensor-core families never store the heatmap: the final 1x1 conv's epilogue reduces each 128-pixel x 64-bin tile to
    an online-softmax partial (conv_tc.cu) and only the merge kernel runs afterwards. Against the same network with
    that fusion switched off (logits written as fp32 NCHW, then the single-pass soft-argmax kernel) the integral
    keypoints agree to summation order."""
    from hrp_b200.model import HoliRobPoseB200
    _, sd = helpers.oracle_for(robot, backbone)
    img, K, kv = helpers.inputs(3, 515)
    outs = []
    for off in ("0", "1"):
        monkeypatch.setenv("HRP_NO_SA_FUSION", off)
        m = HoliRobPoseB200(robot, {"backbone_name": backbone}, device=dev, precision="bf16")
        m.load_state_dict(sd)
        outs.append({k: v.clone() for k, v in m.forward_dict(img.to(dev), K.to(dev), kv.to(dev)).items()})
        n_launch = m.launch_count()
        del m
    assert n_launch > 0
    fused, plain = outs
    assert helpers.maxdiff(fused["uvd"], plain["uvd"]) < 2e-6
    assert helpers.maxdiff(fused["kp2d_int"], plain["kp2d_int"]) < 2e-3
    assert helpers.maxdiff(fused["kp3d_int"], plain["kp3d_int"]) < 1e-5
    for k in ("joint_angles", "rot6d", "root_depth"):          # untouched by the fusion
        assert torch.equal(fused[k], plain[k]), k

def test_host_pipeline_matches_forward_dict(dev):
    """HostPipeline (double-buffered host->device upload overlapping the forward) returns, batch after batch, exactly what
    forward_dict returns for the same inputs."""
    from hrp_b200.model import HostPipeline
    m = gpu_model("panda", "resnet50", dev, "bf16")
    B = 5
    pipe = HostPipeline(m, B)
    batches = [tuple(t.pin_memory() for t in helpers.inputs(B, 300 + i)) for i in range(5)]
    tickets, got = [], []
    for img, K, kv in batches:
        tickets.append(pipe.submit(img, K, kv))
        if len(tickets) > 1:
            got.append({k: v.clone() for k, v in pipe.result(tickets.pop(0)).items()})
    got.append({k: v.clone() for k, v in pipe.result(tickets.pop(0)).items()})
    for (img, K, kv), g in zip(batches, got):
        ref = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
        for k in ref:
            assert torch.equal(ref[k].cpu(), g[k]), k

# ---------------------------------------------------------------------------------------------------- the reference's call sites
def test_reference_evaluator_call_sites(dev):
    """The exact calls of the reference's evaluators on the drop-in: scripts/test.py:157-162 (`model.float(); model.to(device);
    DataParallel(model, ...)`, `test_fps=True`, NINE values unpacked, `times` = (time_root, time_other, time_whole) as
    full_net.py:459-460 returns them) and scripts/real_test.py:291-296 (`test_fps=False`, eight values)."""
    g = helpers.load_golden("fullnet_panda_resnet50.npz")
    _, seed, B = (int(v) for v in g["meta"])
    model = gpu_model("panda", "resnet50", dev)
    img, K, kv = helpers.inputs(B, seed)
    reg_images, root_images, k_values, other_K = img.to(dev), img.clone().to(dev), kv.to(dev), K.to(dev)
    device, device_id = dev, [0]
    model.float()
    model.to(device)
    wrapped = torch.nn.DataParallel(model, device_ids=device_id, output_device=device_id[0])
    pred_pose, pred_rot, pred_trans, pred_root_uv, pred_root_depth, \
        pred_uvd, pred_keypoints3d_int, pred_keypoints3d_fk, times = wrapped(reg_images, root_images, k_values, K=other_K, test_fps=True)
    time_root, time_other, time_whole = times
    assert 0 < time_root < time_whole and 0 < time_other < time_whole and abs(time_root + time_other - time_whole) < 1e-4
    got = dict(joint_angles=pred_pose, rot6d=pred_rot, trans=pred_trans, root_uv=pred_root_uv, root_depth=pred_root_depth,
               uvd=pred_uvd, kp3d_int=pred_keypoints3d_int, kp3d_fk=pred_keypoints3d_fk)
    for k, v in got.items():
        assert helpers.maxdiff(v, g[k]) < (helpers.TOL_PX if k == "root_uv" else 1e-4), k
    eight = wrapped(reg_images, root_images, k_values, K=other_K, test_fps=False)
    assert len(eight) == 8
    for a, b in zip(eight, (pred_pose, pred_rot, pred_trans, pred_root_uv, pred_root_depth, pred_uvd, pred_keypoints3d_int, pred_keypoints3d_fk)):
        assert helpers.maxdiff(a, b) < 1e-5       # graph replay vs the timed un-graphed forward: same kernels


def test_checkpoint_ingestion_and_init_overrides(dev, tmp_path):
    """N3: `load_checkpoint` reads what `save_checkpoint` writes (utils.py:248-254: a dict with 'model_state_dict', keys
    `module.`-prefixed when saved from DataParallel), merges a DepthNet pre-training checkpoint with the factory's re-key
    (`backbone.` -> `rootnet_backbone.`, strict=False: full_net.py:486-500) -- against a golden the reference produced through
    exactly that code path, on un-damped weights with their own BN statistics. Then init_pose / init_rot (full_net.py:268-272)
    and every iterate of the collapsed refinement heads against the oracle's layer-by-layer loop."""
    from hrp_b200.model import HoliRobPoseB200
    g = helpers.load_golden("fullnet_panda_resnet50_ckpt.npz")
    _, seed, B = (int(v) for v in g["meta"])
    ck_main, ck_pre, merged = helpers.checkpoint_case()
    torch.save(ck_main, tmp_path / "curr_best_auc(add)_model.pk")
    torch.save(ck_pre, tmp_path / "depthnet.pk")
    m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision="fp32")
    with pytest.raises(RuntimeError, match="missing"):
        m.load_checkpoint(str(tmp_path / "curr_best_auc(add)_model.pk"))          # strict: the DepthNet tensors are absent
    m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision="fp32")
    missing, unexpected = m.load_checkpoint(str(tmp_path / "curr_best_auc(add)_model.pk"), pretrained_rootnet=str(tmp_path / "depthnet.pk"))
    assert not missing and not unexpected
    img, K, kv = helpers.inputs(B, seed)
    img, K, kv = img.to(dev), K.to(dev), kv.to(dev)
    out = m.forward_dict(img, K, kv)
    d = check_gates(out, g)
    assert d["joint_angles"] < 2e-4 and d["root_depth"] < 2e-4, d
    # two-step load, as the reference's factory + evaluator do it: pretrained DepthNet first (strict=False), then the rest
    m2 = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision="fp32")
    rekeyed = {(k.replace("backbone", "rootnet_backbone") if k.startswith("backbone") else k): v for k, v in ck_pre["model_state_dict"].items()}
    miss, unexp = m2.load_state_dict(rekeyed, strict=False)
    assert miss and unexp == ["xy_layer.weight"]
    with pytest.raises(RuntimeError, match="missing"):
        m2.forward_dict(img, K, kv)
    miss, _ = m2.load_state_dict(ck_main["model_state_dict"], strict=False)
    assert not miss
    out2 = m2.forward_dict(img, K, kv)
    for k in out:
        assert torch.equal(out[k], out2[k]), k
    # init overrides, 8-tuple interface
    ip, ir = torch.from_numpy(g["init_pose"]).to(dev), torch.from_numpy(g["init_rot"]).to(dev)
    o = m(img, img, kv, K, init_pose=ip, init_rot=ir)
    assert helpers.maxdiff(o[0], g["ovr_joint_angles"]) < 1e-4 and helpers.maxdiff(o[1], g["ovr_rot6d"]) < 1e-4
    assert helpers.maxdiff(o[7], g["ovr_kp3d_fk"]) < 1e-4
    assert helpers.maxdiff(o[0], out["joint_angles"]) > 1e-2
    o_pose_only = m(img, img, kv, K, init_pose=ip)                             # one override, the other from the buffer
    assert helpers.maxdiff(o_pose_only[0], g["ovr_joint_angles"]) < 1e-4 and helpers.maxdiff(o_pose_only[1], out["rot6d"]) < 1e-6
    back = m(img, img, kv, K)                                                    # and the defaults return
    assert torch.equal(back[0], out["joint_angles"]) and torch.equal(back[1], out["rot6d"])
    t9 = m(img, img, kv, K, init_pose=ip, init_rot=ir, test_fps=True)          # un-graphed path takes the overrides too
    assert len(t9) == 9 and helpers.maxdiff(t9[0], o[0]) < 1e-6
    # every iterate of both heads (composed affine maps) against the oracle's loop
    from oracle import model as omodel
    om = omodel.OracleModel("panda", merged, open(consts.urdf_path("panda")).read(), "resnet50")
    trace = {}
    om.forward(img.cpu(), img.cpu(), kv.cpu(), K.cpu(), trace=trace, init_pose=ip.cpu(), init_rot=ir.cpu())
    m(img, img, kv, K, init_pose=ip, init_rot=ir)
    iters = m.debug_tensor("head_iters", B).view(B, 4, m.dof + 6)
    for n in range(4):
        assert helpers.maxdiff(iters[:, n, :m.dof], trace["pose_iters"][n]) < 1e-4, n
        assert helpers.maxdiff(iters[:, n, m.dof:], trace["rot_iters"][n]) < 1e-4, n


def test_plan_cache_is_bounded_and_releasable(dev):
    """A caller with ragged batch sizes must not grow device memory by one workspace per size (the cache keeps the
    `max_cached_batches` most recently used sizes); a single stream only ever allocates one plan per size; release_plans
    returns the memory; results are unaffected by eviction."""
    from hrp_b200.model import HoliRobPoseB200
    _, sd = helpers.oracle_for("panda", "resnet50")
    m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision="bf16")
    m.load_state_dict(sd)
    m.set_option("max_cached_batches", 2)
    img, K, kv = helpers.inputs(8, 99)
    img, K, kv = img.to(dev), K.to(dev), kv.to(dev)
    first = {k: v.clone() for k, v in m.forward_dict(img[:8], K[:8], kv[:8]).items()}
    torch.cuda.synchronize()
    ws8 = capi_ws(m, 8)
    free0 = torch.cuda.mem_get_info(dev)[0]
    for b in (7, 6, 5, 4, 3, 8, 7, 6):
        m.forward_dict(img[:b], K[:b], kv[:b])
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info(dev)[0]
    assert free0 - free1 < 1.5 * ws8, (free0 - free1, ws8)          # at most one more plan than before, never 8 of them
    again = m.forward_dict(img[:8], K[:8], kv[:8])
    for k in first:
        assert torch.equal(first[k], again[k]), k
    m.release_plans()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info(dev)[0] >= free0 + ws8 // 2
    again = m.forward_dict(img[:8], K[:8], kv[:8])
    for k in first:
        assert torch.equal(first[k], again[k]), k


def capi_ws(m, B):
    from hrp_b200 import capi
    return int(capi.lib().hrp_workspace_bytes(m._h, B))


# ---------------------------------------------------------------------------------------------------- input side (8f N2)
def test_crop_resize_against_reference_golden(dev):
    """hrp_crop_resize_u8 against what the reference's own data-preparation functions produced on the same frames
    (tests/golden/preprocess.npz): the no-resize branch is bit-exact; resampled crops differ from ATen's CPU bilinear by at
    most one uint8 level on a handful of pixels (fused multiply-adds in ATen's vectorised loop), K' and k_value to fp32
    rounding."""
    from hrp_b200.model import crop_resize
    g = helpers.load_golden("preprocess.npz")
    seed, n = (int(v) for v in g["meta"])
    frames, crop, kbox, K = synth.make_frames(n, seed)
    crops, Kn, kv = crop_resize(cu(frames, dev), cu(crop, dev), cu(K, dev), cu(kbox, dev))
    assert crops.dtype == torch.uint8 and crops.shape == (n, 3, 256, 256)
    d = (crops.cpu().numpy().astype(np.int32) - g["crops"].astype(np.int32))
    assert np.array_equal(crops[0].cpu().numpy(), g["crops"][0])                      # 256 x 256 box: copied, not resampled
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 2e-3, (np.abs(d).max(), (d != 0).mean())
    assert helpers.maxdiff(Kn, g["K"]) < 2e-4
    np.testing.assert_allclose(kv.cpu().numpy(), g["k_value"], rtol=2e-6)
    c2, K2, none = crop_resize(cu(frames, dev), cu(crop, dev), cu(K, dev))            # without the strict box
    assert none is None and torch.equal(c2, crops) and torch.equal(K2, Kn)
    # more boxes than the fixture holds, against the oracle port (which the CPU suite pins to the fixture)
    from oracle import preprocess
    frames, crop, kbox, K = synth.make_frames(12, 7)
    crops, Kn, kv = crop_resize(cu(frames, dev), cu(crop, dev), cu(K, dev), cu(kbox, dev))
    for i in range(12):
        c, Kr, kr = preprocess.crop_resize_one(frames[i], crop[i], K[i], kbox[i])
        dd = crops[i].cpu().numpy().astype(np.int32) - c.astype(np.int32)
        assert np.abs(dd).max() <= 1 and (dd != 0).mean() < 2e-3, i
        assert helpers.maxdiff(Kn[i], Kr) < 2e-4 and abs(float(kv[i]) - float(kr)) < 2e-6 * float(kr)
    e = crop_resize(cu(frames[:0], dev), cu(crop[:0], dev), cu(K[:0], dev))
    assert e[0].shape == (0, 3, 256, 256)


@pytest.mark.parametrize("prec", ["fp32", "bf16", "tf32", "f16"])
def test_forward_from_uint8_crops_equals_forward_on_divided_floats(prec, dev):
    """hrp_forward_u8: the DataLoader's uint8 crops with the `/ 255.` of scripts/test.py:93-96 folded into the stem's input
    pack -- bit-identical to the float path on x = u8 / 255, through forward_dict and through HostPipeline; and the whole
    input side chained (frames -> crop_resize -> forward) against the oracle end to end."""
    from hrp_b200.model import HostPipeline, crop_resize
    m = gpu_model("panda", "resnet50", dev, prec)
    frames, crop, kbox, K = synth.make_frames(5, 23)
    crops, Kn, kv = crop_resize(cu(frames, dev), cu(crop, dev), cu(K, dev), cu(kbox, dev))
    a = m.forward_dict(crops, Kn, kv)
    # IEEE division like the kernel (and the CPU reference); torch's CUDA `x / 255.` multiplies by the reciprocal instead
    xf = torch.from_numpy(crops.cpu().numpy().astype(np.float32) / np.float32(255.0)).to(dev)
    b = m.forward_dict(xf, Kn, kv)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    a2 = m.forward_record(crops, crops.clone(), kv, Kn)[0]                            # distinct reg / root buffers
    assert torch.equal(a2, m.forward_record(xf, xf.clone(), kv, Kn)[0])
    pipe = HostPipeline(m, 5)
    t = pipe.submit(crops.cpu().pin_memory(), Kn.cpu().pin_memory(), kv.cpu().pin_memory())
    got = pipe.result(t)
    for k in a:
        assert torch.equal(a[k].cpu(), got[k]), k
    with pytest.raises(ValueError, match="uint8"):
        m(crops, crops, kv, Kn, test_fps=True)
    if prec == "fp32":
        from oracle import preprocess
        om, _ = helpers.oracle_for("panda", "resnet50")
        ref_in = [preprocess.crop_resize_one(frames[i], crop[i], K[i], kbox[i]) for i in range(5)]
        x = torch.from_numpy(np.stack([r[0] for r in ref_in])).float() / 255.0
        ref = om.forward_dict(x, x, torch.from_numpy(np.stack([r[2] for r in ref_in])), torch.from_numpy(np.stack([r[1] for r in ref_in])))
        check_gates(a, ref)


# ---------------------------------------------------------------------------------------------------- families at config scale
@pytest.mark.parametrize("prec", ["tf32", "bf16", "f16"])
@pytest.mark.parametrize("robot,B", [("panda", 64), ("kuka", 256), ("baxter", 128)])
def test_tensor_core_families_against_oracle_at_config_batch(prec, robot, B, dev):
    """BASELINE configs 2-4 at their per-GPU batch sizes (which switch kernel paths: branch chains, lane shares, slab cost
    model): sampled frames of the big batch against the ORACLE (not against the library itself), each family at its stated
    tolerance."""
    m = gpu_model(robot, "resnet50", dev, prec)
    om, _ = helpers.oracle_for(robot, "resnet50")
    img, K, kv = helpers.inputs(B, 4000 + B)
    big = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    idx = [0, B // 2 + 1, B - 1]
    ref = om.forward_dict(img[idx], img[idx], kv[idx], K[idx])
    t = FAMILY_TOL[prec]
    d = {k: helpers.maxdiff(big[k][idx], ref[k]) for k in ref}
    print(prec, robot, B, {k: "%.2e" % v for k, v in d.items()})
    assert d["joint_angles"] < t["joint_angles"] and d["root_depth"] < t["root_depth"] and d["rot6d"] < t["rot6d"], d
    assert max(d["kp2d_int"], d["kp2d_fk"], d["root_uv"]) < t["px"] and d["uvd"] < t["uvd"], d
    assert max(d["kp3d_int"], d["kp3d_fk"], d["trans"]) < t["m3d"], d
    m.release_plans()


@pytest.mark.parametrize("robot,backbone", helpers.UNDAMPED_CASES)
def test_families_on_undamped_weights_reported(robot, backbone, dev):
    """SURVEY 8d's weight recipe as written (unit-gain residual branches): the fp32 family is held to the gates; the
    tensor-core families' deviations are REPORTED (gpurun_out/r2_undamped_deviation.jsonl, summarised in DESIGN.md section 2),
    not gated -- this network amplifies operand rounding chaotically, in the reference's own forward too."""
    import json, os
    from hrp_b200.model import HoliRobPoseB200
    g = helpers.load_golden("fullnet_%s_%s_undamped.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    _, sd = helpers.oracle_for(robot, backbone, wseed, recipe="undamped")
    img, K, kv = helpers.inputs(B, seed)
    names = ["joint_angles", "rot6d", "root_depth", "root_uv", "uvd", "kp3d_int", "kp3d_fk", "kp2d_int", "kp2d_fk"]
    for prec in ("fp32", "tf32", "f16", "bf16"):
        m = HoliRobPoseB200(robot, {"backbone_name": backbone}, device=dev, precision=prec)
        m.load_state_dict(sd)
        out = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
        d = {k: helpers.maxdiff(out[k], g[k]) for k in names}
        print("undamped", prec, robot, backbone, {k: "%.2e" % v for k, v in d.items()})
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "r2_undamped_deviation.jsonl"), "a") as f:
            f.write(json.dumps(dict(recipe="undamped", precision=prec, robot=robot, backbone=backbone, maxdiff=d)) + "\n")
        if prec == "fp32":
            check_gates(out, g)
        assert all(np.isfinite(v) for v in d.values())
        del m


# Tensor-core families against the reference's fp32 forward (goldens). The TF32 family is held to the north_star parity
# gates themselves -- 1e-3 rad, 1 mm, 0.5 px -- on EVERY configuration: single-pass TF32 (operands rounded to nearest, fp32
# accumulation in TMEM) on the ResNet-50 keypoint backbone and the DepthNet (measured worst case 5.4e-4 rad / 0.2 mm /
# 0.15 px), 3xTF32 on the layers of an HRNet-W32 keypoint backbone (single-pass reaches 1.8e-3 rad there: twice as many
# sequential roundings in front of the joint-angle heads). bf16 carries the separately stated tolerance 2e-2 rad / 5 mm /
# 3 px (measured worst case 1.4e-2 rad, 0.9 mm, 2.1 px). The fp32 family meets the gates everywhere
# (test_fullnet_against_reference_golden). DESIGN.md section 2.
FAMILY_TOL = {
    "tf32": dict(joint_angles=helpers.TOL_RAD, root_depth=helpers.TOL_DEPTH_M, px=helpers.TOL_PX, rot6d=2e-3, uvd=1e-3, m3d=2e-3, rel=0.01),
    # IEEE-half operands and activations (11-bit significand like TF32): held to the same north_star gates
    "f16": dict(joint_angles=helpers.TOL_RAD, root_depth=helpers.TOL_DEPTH_M, px=helpers.TOL_PX, rot6d=2e-3, uvd=1e-3, m3d=2e-3, rel=0.01),
    "tf32x3": dict(joint_angles=2e-4, root_depth=2e-4, px=0.1, rot6d=2e-4, uvd=2e-4, m3d=5e-4, rel=1e-3),
    "bf16": dict(joint_angles=2e-2, root_depth=5e-3, px=3.0, rot6d=2e-2, uvd=5e-3, m3d=1.5e-2, rel=0.05),
}


@pytest.mark.parametrize("robot,backbone", [("panda", "resnet50"), ("baxter", "hrnet32")])
def test_fullnet_3xtf32_everywhere_is_fp32_grade(robot, backbone, dev):
    """precision="tf32x3": every conv layer of both backbones as 3xTF32 -- the safety net for the whole conv engine
    (SURVEY 7.3 H3c); as close to the reference goldens as the fp32 FFMA family."""
    g = helpers.load_golden("fullnet_%s_%s.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    m = gpu_model(robot, backbone, dev, "tf32x3")
    img, K, kv = helpers.inputs(B, seed)
    out = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    t = FAMILY_TOL["tf32x3"]
    d = {k: helpers.maxdiff(out[k], g[k]) for k in out}
    print("tf32x3", robot, backbone, {k: "%.2e" % v for k, v in d.items()})
    assert d["joint_angles"] < t["joint_angles"] and d["root_depth"] < t["root_depth"] and d["rot6d"] < t["rot6d"], d
    assert max(d["kp2d_int"], d["kp2d_fk"], d["root_uv"]) < t["px"] and d["uvd"] < t["uvd"], d
    del _models[(robot, backbone, "tf32x3")]


@pytest.mark.parametrize("prec", ["tf32", "bf16", "f16"])
@pytest.mark.parametrize("robot,backbone", helpers.FULLNET_CASES)
def test_fullnet_tensor_core_families_against_reference_golden(prec, robot, backbone, dev):
    g = helpers.load_golden("fullnet_%s_%s.npz" % (robot, backbone))
    wseed, seed, B = (int(v) for v in g["meta"])
    m = gpu_model(robot, backbone, dev, prec)
    img, K, kv = helpers.inputs(B, seed)
    out = m.forward_dict(img.to(dev), K.to(dev), kv.to(dev))
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk", "kp2d_int", "kp2d_fk"]
    d = {k: helpers.maxdiff(out[k], g[k]) for k in names}
    print(prec, robot, backbone, {k: "%.2e" % v for k, v in d.items()})
    t = FAMILY_TOL[prec]
    if prec == "f16" and backbone == "hrnet32":
        # the f16 family has no hi/lo split: with the HRNet-W32 keypoint backbone it behaves like single-pass TF32 did
        # (~2e-3 rad: twice as many sequential 11-bit roundings in front of the heads). The PARITY mode for that variant is
        # tf32 (3xTF32 there), which meets the gates; f16's stated tolerance on it is 3e-3 rad / 3e-3 rot6d, gates elsewhere.
        t = dict(t, joint_angles=3e-3, rot6d=3e-3)
    assert d["joint_angles"] < t["joint_angles"] and d["root_depth"] < t["root_depth"] and d["rot6d"] < t["rot6d"], d
    assert max(d["kp2d_int"], d["kp2d_fk"], d["root_uv"]) < t["px"] and d["uvd"] < t["uvd"], d
    assert max(d["kp3d_int"], d["kp3d_fk"], d["trans"]) < t["m3d"], d
    # the 2048-d features feeding the heads, relative RMS
    for name, key in (("xf", "probe_xf"), ("img_feat", "probe_img_feat")):
        a = m.debug_tensor(name, B).view(B, -1).cpu().double().numpy()
        rel = float(np.linalg.norm(a - g[key]) / np.linalg.norm(g[key]))
        assert rel < t["rel"], (name, rel)
    assert m.launch_count() > 300


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["panda", "kuka", "baxter"])
def test_metrics_tail_against_reference_golden(robot, dev):
    """compute_metrics_batch / summary_add_pck on the device (8f N4) against the reference's own outputs and the oracle port."""
    from hrp_b200 import metrics as hm
    from hrp_b200.model import FkRobot
    from oracle import metrics as ometrics
    g = helpers.load_golden("metrics_%s.npz" % robot)
    seed, n, root = (int(v) for v in g["meta"])
    d = synth.make_metrics_inputs(robot, n, seed)
    fk = FkRobot(robot)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    names = ["error3d", "error2d", "dis3d", "dis2d", "l1_jointerror", "mean_jointerror", "error_depth", "batch_error_relative", "error3d_relative"]
    # the reference's call (scripts/test.py:167-181): poses in, FK + projection with the original camera inside
    res = hm.compute_metrics_batch(fk, T(g["gt_xyz"]), T(g["gt_uv"]), T(d["K"]), T(d["gt_q"]), pred_joint=T(d["q"]), pred_rot=T(d["rot"]),
                                   pred_trans=T(d["trans"]), pred_depth=None, pred_xy=None, pred_xyz_integral=None, reference_keypoint_id=root)
    assert len(res) == 9
    # FK runs in fp32 on both sides with different operation orders: 2e-6 m on the keypoints, amplified by the projection (f/z ~ 1e3)
    tol = dict(error3d=5e-6, error2d=5e-3, dis3d=5e-6, dis2d=5e-3, l1_jointerror=1e-6, mean_jointerror=1e-6, error_depth=5e-6,
               batch_error_relative=5e-6, error3d_relative=5e-6)
    for k, v in zip(names, res):
        np.testing.assert_allclose(np.asarray(v, np.float32), g[k], rtol=2e-5, atol=tol[k], equal_nan=True, err_msg=k)
    assert np.isnan(res[1][3])                                                       # no keypoint inside the frame: 0/0 like numpy
    # given 3-D points instead of a pose (metrics.py:22-26): no FK, joint errors reported as zeros
    res2 = hm.compute_metrics_batch(fk, T(g["gt_xyz"]), T(g["gt_uv"]), T(d["K"]), T(d["gt_q"]), pred_joint=None, pred_rot=None, pred_trans=None,
                                    pred_xyz_integral=T(g["pred_xyz"]), reference_keypoint_id=root)
    want = ometrics.batch_errors(g["pred_xyz"], g["gt_xyz"], g["gt_uv"], d["K"], d["gt_q"], None, root, robot)
    for k, v, w in zip(names, res2, want):
        np.testing.assert_allclose(np.asarray(v, np.float32), g["nojoint_" + k], rtol=2e-5, atol=tol[k] if "2d" in k else 2e-6, equal_nan=True, err_msg=k)
        np.testing.assert_allclose(np.asarray(v, np.float32), np.asarray(w, np.float32), rtol=2e-5, atol=tol[k] if "2d" in k else 2e-6, equal_nan=True, err_msg=k)
    # the run summary from the SAME per-frame lists the reference summarised (exact inputs -> tight tolerance)
    ok = ~np.isnan(g["error2d"])
    s = hm.summary_add_pck({"dis3d": list(g["error3d"][ok]), "dis2d": list(g["error2d"][ok])})
    assert list(s.keys()) == [str(k) for k in g["summary_keys"]]
    for k, v in zip(g["summary_keys"], g["summary_values"]):
        np.testing.assert_allclose(float(s[str(k)]), v, rtol=1e-6, atol=1e-9, err_msg=str(k))
    # device-resident accumulation over batches (ErrorLog) gives the same summary; odd and even list lengths for the median
    log = hm.ErrorLog(dev, capacity=64)
    pf, _, _, _ = hm.metrics_batch_device(fk, T(g["gt_xyz"]), T(g["gt_uv"]), T(d["K"]), T(d["gt_q"]), pred_xyz_integral=T(g["pred_xyz"]))
    keep = torch.from_numpy(ok).to(dev)
    for lo in range(0, n, 96):
        sel = pf[lo:lo + 96][keep[lo:lo + 96]]
        log.extend(sel)
    assert log.n == int(ok.sum())
    s2 = log.summary()
    for k in s:
        np.testing.assert_allclose(float(s2[k]), float(s[k]), rtol=1e-5, atol=1e-7, err_msg=k)
    for m in (log.n - 1, 7):
        e3, e2 = g["error3d"][ok][:m], g["error2d"][ok][:m]
        s3 = hm.summary_add_pck({"dis3d": T(e3), "dis2d": T(e2)})
        w3 = ometrics.summary(e3, e2)
        for k in w3:
            np.testing.assert_allclose(float(s3[k]), float(w3[k]), rtol=1e-6, atol=1e-9, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("name", list(helpers.VARIANT_CASES))
def test_constructor_variants_against_reference_golden(name, prec, dev):
    """8f N4: direct_reg_rot + add_fc + multi_kp, rot_iterative_matmul and reg_joint_map on the device, against the reference built
    with those switches (tests/golden/variant_*.npz) and, at a second batch size, against the oracle port."""
    from hrp_b200.model import HoliRobPoseB200
    from oracle import model as omodel
    g = helpers.load_golden("variant_%s.npz" % name)
    wseed, seed, B = (int(v) for v in g["meta"])
    cfg, ctor = helpers.VARIANT_CASES[name]
    sd = helpers.variant_state_dict(name, wseed)
    model = HoliRobPoseB200("panda", dict(cfg), device=dev, precision=prec)
    model.load_state_dict(sd)
    img, K, kv = helpers.inputs(B, seed)
    out = model(img.to(dev), img.to(dev), kv.to(dev), K.to(dev))
    names = ["joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk"]
    if cfg.get("multi_kp"):
        assert len(out) == 9                                                   # full_net.py:462-464
        names = names[:5] + ["depths"] + names[5:]
        assert tuple(out[5].shape) == (B, len(cfg["kps_need_depth"]))
    else:
        assert len(out) == 8
    res = dict(zip(names, out))
    t_rad, t_m, t_px = (2e-5, 2e-5, 1e-2) if prec == "fp32" else (1e-3, 1e-3, 0.5)
    if name == "jointmap" and prec == "fp32":
        # the joint angle is an expectation over a softmax of high-contrast maps (fixture gain 6) times a ~6 rad range: it
        # amplifies the fp32 summation-order differences of the 53-layer trunk ~10x (measured 1.3e-4 rad); still 8x inside the gate
        t_rad, t_m, t_px = 3e-4, 3e-4, 0.1
    tol = dict(joint_angles=t_rad, rot6d=t_rad, trans=t_m, root_depth=t_m, uvd=t_rad, kp3d_int=t_m, kp3d_fk=t_m, root_uv=t_px, depths=t_m)
    if name == "jointmap" and prec in ("tf32", "f16"):
        # REPORTED, not a parity claim: with 11-bit operands the same amplification puts the joint angles of this fixture 6e-2 rad
        # from the reference (the other outputs stay inside the gates). The parity modes of this variant are fp32 and tf32x3 (below)
        tol.update(joint_angles=0.15, kp3d_fk=0.08)
    for k, v in res.items():
        assert helpers.maxdiff(v, g[k]) < tol[k], (k, helpers.maxdiff(v, g[k]))
    # a batch the golden does not hold, with init_rot overridden (only the refinement variant reads it)
    om = omodel.OracleModel("panda", sd, open(consts.urdf_path("panda")).read(), "resnet50", ctor=ctor)
    if name == "jointmap":
        lo, hi = np.asarray(consts.ROBOTS["panda"]["bounds"], np.float32).T
        q = res["joint_angles"].cpu().numpy()
        assert (q >= lo - 1e-6).all() and (q <= hi + 1e-6).all() and q.std(0).max() > 0.1    # inside the bounds, not mid-range everywhere
    img2, K2, kv2 = helpers.inputs(5, seed + 1)
    r0 = torch.tensor([[0.9, 0.1, -0.2, 0.05, 1.1, 0.3]]).expand(5, 6).contiguous()
    want = om.forward(img2, img2, kv2, K2, init_rot=r0)
    got = model(img2.to(dev), img2.to(dev), kv2.to(dev), K2.to(dev), init_rot=r0.to(dev))
    if cfg.get("multi_kp"):
        got = got[:5] + got[6:]
    for k, a, b in zip(["joint_angles", "rot6d", "trans", "root_uv", "root_depth"], got, want):
        assert helpers.maxdiff(a, b) < tol[k], (k, helpers.maxdiff(a, b))
    if name == "jointmap" and prec == "tf32":
        m3 = HoliRobPoseB200("panda", dict(cfg), device=dev, precision="tf32x3")
        m3.load_state_dict(sd)
        o3 = m3(img.to(dev), img.to(dev), kv.to(dev), K.to(dev))
        assert helpers.maxdiff(o3[0], g["joint_angles"]) < 1e-3 and helpers.maxdiff(o3[7], g["kp3d_fk"]) < 1e-3      # measured 4.0e-4 rad
    if name == "rotmatmul":
        iters = model.debug_tensor("head_iters", 5).reshape(5, model.n_iter, model.dof + 6)
        trace = {}
        om.forward(img2, img2, kv2, K2, init_rot=r0, trace=trace)
        for n in range(model.n_iter):                                          # every composed rotation, not just the last
            assert helpers.maxdiff(iters[:, n, model.dof:], trace["rot_iters"][n]) < t_rad * 2
            R = iters[:, n, model.dof:].reshape(5, 2, 3)
            assert float((R.norm(dim=2) - 1).abs().max()) < 1e-5 and float((R[:, 0] * R[:, 1]).sum(1).abs().max()) < 1e-5


@pytest.mark.gpu
def test_peer_memory_gather_single_rank(dev):
    """hrp_p2p_* with a world of one (the multi-rank protocol is exercised by scripts/p2p_gather_test.py under torchrun: 64
    overlapping steps on 2 and 8 GPUs bit-equal to NCCL's all-gather, profiles/r02_p2p_gather_vs_nccl.jsonl)."""
    import ctypes as C
    from hrp_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    n = 7101                                                        # not a multiple of four floats: the window rounds up to 16 bytes
    padded = (n * 4 + 15) // 16 * 4
    capi.check(L.hrp_p2p_create(0, 1, n * 4, dev.index, C.byref(h)))
    src = torch.zeros(padded, device=dev)
    out = torch.empty(padded, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for step in range(5):
        src[:n] = torch.arange(n, device=dev, dtype=torch.float32) + step
        capi.check(L.hrp_p2p_all_gather(h, C.c_void_p(src.data_ptr()), padded * 4, C.c_void_p(out.data_ptr()), st))
        assert torch.equal(out[:n], src[:n])
    capi.check(L.hrp_p2p_status(h))
    assert L.hrp_p2p_all_gather(h, C.c_void_p(src.data_ptr()), 64, C.c_void_p(out.data_ptr()), st) == -1          # HRP_ERR_INVALID: not the window's record size
    L.hrp_p2p_destroy(h)
