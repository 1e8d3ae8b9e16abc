"""Tensor inventory of the HoliRobPose inference network in the reference's state-dict naming.

One generator yields (name, shape, kind) for every tensor `RootNetwithRegInt.state_dict()` holds
(lib/models/full_net.py:77-176,211-212; backbones/HRnet.py:274-339; backbones/Resnet.py:6-55; hrnet_w32.yaml:50-86).
It is what `load_state_dict` validates against and what the synthetic-weight generator fills; the C++ graph builder
(csrc/network.cu) asks for the same names and the test-suite checks the two lists agree.
"""
from . import consts

HR_STAGES = (  # (modules, blocks per branch, channels per branch) for stage2..4 of HRNet-W32
    (1, 4, (32, 64)),
    (4, 4, (32, 64, 128)),
    (3, 4, (32, 64, 128, 256)),
)
HR_HEAD = (32, 64, 128, 256)   # cls-head bottleneck planes (HRnet.py:343)
RESNET50_BLOCKS = (3, 4, 6, 3)
RESNET50_PLANES = (64, 128, 256, 512)


def _bn(p, c):
    yield p + ".weight", (c,), "bn_w"
    yield p + ".bias", (c,), "bn_b"
    yield p + ".running_mean", (c,), "bn_mean"
    yield p + ".running_var", (c,), "bn_var"
    yield p + ".num_batches_tracked", (), "bn_nbt"


def _conv(p, cin, cout, k, bias=False):
    yield p + ".weight", (cout, cin, k, k), "conv_w"
    if bias:
        yield p + ".bias", (cout,), "conv_b"


def _conv_bn(pc, pb, cin, cout, k, bias=False):
    yield from _conv(pc, cin, cout, k, bias)
    yield from _bn(pb, cout)


def _bottleneck(p, cin, planes, downsample):
    yield from _conv_bn(p + ".conv1", p + ".bn1", cin, planes, 1)
    yield from _conv_bn(p + ".conv2", p + ".bn2", planes, planes, 3)
    yield from _conv_bn(p + ".conv3", p + ".bn3", planes, planes * 4, 1)
    if downsample:
        yield from _conv_bn(p + ".downsample.0", p + ".downsample.1", cin, planes * 4, 1)


def _basic(p, c):
    yield from _conv_bn(p + ".conv1", p + ".bn1", c, c, 3)
    yield from _conv_bn(p + ".conv2", p + ".bn2", c, c, 3)


def resnet50(p):
    yield from _conv_bn(p + "conv1", p + "bn1", 3, 64, 7)
    cin = 64
    for li, (nb, planes) in enumerate(zip(RESNET50_BLOCKS, RESNET50_PLANES)):
        for b in range(nb):
            yield from _bottleneck("%slayer%d.%d" % (p, li + 1, b), cin, planes, b == 0)
            cin = planes * 4


def hrnet_w32(p, heatmap_channels=0):
    yield from _conv_bn(p + "conv1", p + "bn1", 3, 64, 3)
    yield from _conv_bn(p + "conv2", p + "bn2", 64, 64, 3)
    cin = 64
    for b in range(4):
        yield from _bottleneck("%slayer1.%d" % (p, b), cin, 64, b == 0)
        cin = 256
    pre = (256,)
    for si, (nmod, nblk, chans) in enumerate(HR_STAGES):
        t = "%stransition%d" % (p, si + 1)
        for i, c in enumerate(chans):
            if i < len(pre):
                if pre[i] != c:
                    yield from _conv_bn("%s.%d.0" % (t, i), "%s.%d.1" % (t, i), pre[i], c, 3)
            else:  # one new branch per stage: a single stride-2 conv from the last previous branch
                yield from _conv_bn("%s.%d.0.0" % (t, i), "%s.%d.0.1" % (t, i), pre[-1], c, 3)
        for m in range(nmod):
            s = "%sstage%d.%d" % (p, si + 2, m)
            for bi, c in enumerate(chans):
                for k in range(nblk):
                    yield from _basic("%s.branches.%d.%d" % (s, bi, k), c)
            for i, ci in enumerate(chans):
                for j, cj in enumerate(chans):
                    f = "%s.fuse_layers.%d.%d" % (s, i, j)
                    if j > i:
                        yield from _conv_bn(f + ".0", f + ".1", cj, ci, 1)
                    elif j < i:
                        for k in range(i - j):
                            co = ci if k == i - j - 1 else cj
                            yield from _conv_bn("%s.%d.0" % (f, k), "%s.%d.1" % (f, k), cj, co, 3)
        pre = chans
    for i, (c, h) in enumerate(zip(pre, HR_HEAD)):
        yield from _bottleneck("%sincre_modules.%d.0" % (p, i), c, h, True)
    for i in range(3):
        d = "%sdownsamp_modules.%d" % (p, i)
        yield from _conv_bn(d + ".0", d + ".1", HR_HEAD[i] * 4, HR_HEAD[i + 1] * 4, 3, bias=True)
    yield from _conv_bn(p + "final_feat_layer.0", p + "final_feat_layer.1", 1024, 2048, 1, bias=True)
    if heatmap_channels:
        yield from _conv(p + "final_layer", 32, heatmap_channels, 1, bias=True)


def _linear(p, cin, cout):
    yield p + ".weight", (cout, cin), "lin_w"
    yield p + ".bias", (cout,), "lin_b"


VARIANT_DEFAULTS = dict(direct_reg_rot=False, rot_iterative_matmul=False, add_fc=False, depth_num=1, reg_joint_map=False,
                        joint_conv_dim=())


def full_net(robot, backbone="resnet50", variant=None):
    """Ordered tensor list of RootNetwithRegInt(robot, backbone_name=backbone, rootnet_backbone_name='hrnet32').

    variant: the constructor switches that change the tensor list (full_net.py:107-131, 149-176): direct_reg_rot (a seven-layer
    rotation regressor instead of the refinement loop), add_fc (the DepthNet's bottleneck MLP), depth_num (multi_kp: one depth
    output per entry of kps_need_depth), reg_joint_map + joint_conv_dim (joint angles from per-joint maps instead of the
    refinement MLP). rot_iterative_matmul changes arithmetic only."""
    v = dict(VARIANT_DEFAULTS, **(variant or {}))
    spec = consts.ROBOTS[robot]
    dof, nkpt = spec["dof"], spec["nkpt"]
    hm = nkpt * consts.DEPTH_DIM
    if backbone in ("resnet", "resnet50"):
        yield from resnet50("reg_backbone.")
        cin = 2048
        for i in range(3):  # ConvTranspose2d weight is [Cin, Cout, 4, 4] (full_net.py:218-226)
            yield "deconv_layers.%d.weight" % (3 * i), (cin, 256, 4, 4), "deconv_w"
            yield from _bn("deconv_layers.%d" % (3 * i + 1), 256)
            cin = 256
        yield from _conv("final_layer", 256, hm, 1, bias=True)
    elif backbone in ("hrnet", "hrnet32"):
        yield from hrnet_w32("reg_backbone.", hm)
    else:
        raise ValueError("unsupported backbone_name %r (supported: resnet50, hrnet32)" % backbone)
    if v["reg_joint_map"]:                                      # full_net.py:92-101, 240-258 (nn.Sequential indices 0,1 / 3,4 / 6,7)
        cin = consts.FEATURE_DIM
        for i, c in enumerate(v["joint_conv_dim"]):
            yield from _conv_bn("joint_conv_layers.%d" % (3 * i), "joint_conv_layers.%d" % (3 * i + 1), cin, c, 3, bias=True)
            cin = c
        yield from _conv("joint_final_layer", cin, dof, 1, bias=True)
    else:
        yield from _linear("fc_pose_1", consts.FEATURE_DIM + dof, 1024)
        yield from _linear("fc_pose_2", 1024, 1024)
        yield from _linear("decpose", 1024, dof)
    if v["direct_reg_rot"]:                                     # full_net.py:110-117
        yield from _linear("fc_rot_1", consts.FEATURE_DIM, 1024)
        for i in range(2, 7):
            yield from _linear("fc_rot_%d" % i, 1024, 1024)
    else:
        yield from _linear("fc_rot_1", consts.FEATURE_DIM + consts.ROT_DIM, 1024)
        yield from _linear("fc_rot_2", 1024, 1024)
    yield from _linear("decrot", 1024, consts.ROT_DIM)
    yield from hrnet_w32("rootnet_backbone.", 0)
    if v["add_fc"]:                                             # full_net.py:156-163
        yield from _linear("depth_fc_d1", consts.FEATURE_DIM, 1024)
        yield from _linear("depth_fc_d2", 1024, 512)
        yield from _bn("depth_bn", 512)
        yield from _linear("depth_fc_u2", 512, 1024)
        yield from _linear("depth_fc_u1", 1024, consts.FEATURE_DIM)
    yield from _conv("depth_layer", consts.FEATURE_DIM, int(v["depth_num"]), 1, bias=True)
    yield "init_pose", (1, dof), "buf"
    yield "init_rot", (1, consts.ROT_DIM), "buf"


def flops_per_frame(robot, backbone="resnet50"):
    """Algorithmic 2*MAC of convs + linears per frame at 256x256 (SURVEY.md §8d / BASELINE.md §3), in FLOP."""
    hrnet = {0: 23.299e9, 7: 23.416e9, 8: 23.433e9, 17: 23.584e9}
    nk = consts.ROBOTS[robot]["nkpt"]
    mlp = 0.051e9
    if backbone in ("resnet", "resnet50"):
        final = {7: 0.940e9, 8: 1.074e9, 17: 2.282e9}[nk]
        return hrnet[0] + 10.677e9 + 3.758e9 + final + mlp
    return hrnet[0] + hrnet[nk] + mlp
