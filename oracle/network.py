"""Oracle: backbones and heads as pure functions over a reference-format state dict (fp32, PyTorch CPU).

Follows lib/models/backbones/HRnet.py:247-265,499-570 (HRNet-W32 trunk, fusion, cls head),
lib/models/backbones/Resnet.py:57-68,100-139 (ResNet-50 trunk), lib/models/full_net.py:214-238,289-355,376-444
(deconv head, DepthNet head, iterative linear heads). `Calib` switches BatchNorm to batch statistics and records them
(used once by scripts/make_bn_calib.py to produce the calibrated synthetic weights).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default; no module in the reference overrides it


class Calib:
    """When passed as `calib`, BN uses batch statistics and stores them (momentum=None semantics, one batch)."""

    def __init__(self):
        self.stats = {}


def bn(x, sd, p, calib=None):
    if calib is not None:
        mean = x.mean((0, 2, 3))
        var_b = x.var((0, 2, 3), unbiased=False)
        n = x.numel() / x.shape[1]
        calib.stats[p + ".running_mean"] = mean.clone()
        calib.stats[p + ".running_var"] = (var_b * n / (n - 1)).clone()   # running_var is the unbiased estimate
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, EPS)


def conv(x, sd, p, stride=1, pad=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride, pad)


def cbr(x, sd, pc, pb, stride=1, pad=0, relu=True, calib=None):
    y = bn(conv(x, sd, pc, stride, pad), sd, pb, calib)
    return F.relu(y) if relu else y


def bottleneck(x, sd, p, stride=1, calib=None):
    # Resnet.py:121-139 / HRnet.py:78-98 (stride sits on the 3x3)
    y = cbr(x, sd, p + ".conv1", p + ".bn1", calib=calib)
    y = cbr(y, sd, p + ".conv2", p + ".bn2", stride, 1, calib=calib)
    y = cbr(y, sd, p + ".conv3", p + ".bn3", relu=False, calib=calib)
    if (p + ".downsample.0.weight") in sd:
        x = cbr(x, sd, p + ".downsample.0", p + ".downsample.1", stride, 0, relu=False, calib=calib)
    return F.relu(y + x)


def basic(x, sd, p, calib=None):
    # HRnet.py:41-57
    y = cbr(x, sd, p + ".conv1", p + ".bn1", 1, 1, calib=calib)
    y = cbr(y, sd, p + ".conv2", p + ".bn2", 1, 1, relu=False, calib=calib)
    return F.relu(y + x)


def resnet50(x, sd, p, calib=None):
    x = cbr(x, sd, p + "conv1", p + "bn1", 2, 3, calib=calib)
    x = F.max_pool2d(x, 3, 2, 1)
    for li, nb in enumerate((3, 4, 6, 3)):
        for b in range(nb):
            x = bottleneck(x, sd, "%slayer%d.%d" % (p, li + 1, b), 2 if (b == 0 and li > 0) else 1, calib)
    return x


def hr_module(xs, sd, p, calib=None):
    # HRnet.py:247-265
    n = len(xs)
    xs = list(xs)
    for i in range(n):
        for k in range(4):
            xs[i] = basic(xs[i], sd, "%s.branches.%d.%d" % (p, i, k), calib)
    out = []
    for i in range(n):
        y = None
        for j in range(n):
            f = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = cbr(xs[j], sd, f + ".0", f + ".1", relu=False, calib=calib)
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = cbr(t, sd, "%s.%d.0" % (f, k), "%s.%d.1" % (f, k), 2, 1, relu=(k != i - j - 1), calib=calib)
            y = t if y is None else y + t
        out.append(F.relu(y))
    return out


def hrnet_w32(x, sd, p, heatmap=False, calib=None):
    """Returns (heatmap logits or None, feat [B,2048])."""
    x = cbr(x, sd, p + "conv1", p + "bn1", 2, 1, calib=calib)
    x = cbr(x, sd, p + "conv2", p + "bn2", 2, 1, calib=calib)
    for b in range(4):
        x = bottleneck(x, sd, "%slayer1.%d" % (p, b), 1, calib)
    t = p + "transition1"
    ys = [cbr(x, sd, t + ".0.0", t + ".0.1", 1, 1, calib=calib), cbr(x, sd, t + ".1.0.0", t + ".1.0.1", 2, 1, calib=calib)]
    ys = hr_module(ys, sd, p + "stage2.0", calib)
    for si, nmod in ((3, 4), (4, 3)):
        t = "%stransition%d.%d.0" % (p, si - 1, si - 1)
        ys = ys + [cbr(ys[-1], sd, t + ".0", t + ".1", 2, 1, calib=calib)]   # new branch reads y_list[-1], HRnet.py:519
        for m in range(nmod):
            ys = hr_module(ys, sd, "%sstage%d.%d" % (p, si, m), calib)
    hm = conv(ys[0], sd, p + "final_layer") if heatmap else None
    # classification head, HRnet.py:537-548
    y = bottleneck(ys[0], sd, p + "incre_modules.0.0", 1, calib)
    for i in range(3):
        d = "%sdownsamp_modules.%d" % (p, i)
        y = bottleneck(ys[i + 1], sd, "%sincre_modules.%d.0" % (p, i + 1), 1, calib) + \
            cbr(y, sd, d + ".0", d + ".1", 2, 1, calib=calib)
    y = cbr(y, sd, p + "final_feat_layer.0", p + "final_feat_layer.1", calib=calib)
    return hm, y.mean((2, 3))


def deconv_head(x, sd, calib=None):
    # full_net.py:214-238, 353-355
    for i in range(3):
        x = F.conv_transpose2d(x, sd["deconv_layers.%d.weight" % (3 * i)], None, 2, 1)
        x = F.relu(bn(x, sd, "deconv_layers.%d" % (3 * i + 1), calib))
    return conv(x, sd, "final_layer")


def iterative_head(xf, state, sd, fc1, fc2, dec, n_iter=4, trace=None):
    # full_net.py:381-394 / 431-444: purely linear refinement, dropout is the identity in eval mode
    for _ in range(n_iter):
        xc = torch.cat([xf, state], 1)
        xc = F.linear(xc, sd[fc1 + ".weight"], sd[fc1 + ".bias"])
        xc = F.linear(xc, sd[fc2 + ".weight"], sd[fc2 + ".bias"])
        state = F.linear(xc, sd[dec + ".weight"], sd[dec + ".bias"]) + state
        if trace is not None:
            trace.append(state.clone())
    return state


def rot6d_to_rotmat(r):
    # lib/utils/geometries.py:100-115: rows (x, y, z) with x = a1/|a1|, z = (x X a2)/|.|, y = z X x
    x = r[..., 0:3] / torch.norm(r[..., 0:3], p=2, dim=-1, keepdim=True)
    z = torch.cross(x, r[..., 3:6], dim=-1)
    z = z / torch.norm(z, p=2, dim=-1, keepdim=True)
    y = torch.cross(z, x, dim=-1)
    return torch.stack((x, y, z), -2)


def matmul_rot_head(xf, state, sd, n_iter=4, trace=None):
    # full_net.py:413-429 (rot_iterative_matmul): the regressed 6-vector is a rotation COMPOSED with the current one,
    # and the state is the first two rows of the product (geometries.py:117-132)
    for _ in range(n_iter):
        xc = torch.cat([xf, state], 1)
        xc = F.linear(xc, sd["fc_rot_1.weight"], sd["fc_rot_1.bias"])
        xc = F.linear(xc, sd["fc_rot_2.weight"], sd["fc_rot_2.bias"])
        d = F.linear(xc, sd["decrot.weight"], sd["decrot.bias"])
        state = (rot6d_to_rotmat(d) @ rot6d_to_rotmat(state))[..., :2, :].reshape(-1, 6)
        if trace is not None:
            trace.append(state.clone())
    return state


def direct_rot_head(xf, sd):
    # full_net.py:395-409 (direct_reg_rot): seven linear layers, fc_rot_1's output added back before the decoder
    x1 = F.linear(xf, sd["fc_rot_1.weight"], sd["fc_rot_1.bias"])
    x = x1
    for i in range(2, 7):
        x = F.linear(x, sd["fc_rot_%d.weight" % i], sd["fc_rot_%d.bias" % i])
    return F.linear(x + x1, sd["decrot.weight"], sd["decrot.bias"])


def depth_add_fc(feat, sd):
    # full_net.py:296-313 (add_fc): 2048 -> 1024 -> 512 -> BatchNorm1d -> LeakyReLU -> 1024 (+skip)/2 -> 2048 (+skip)/2
    f1 = F.linear(feat, sd["depth_fc_d1.weight"], sd["depth_fc_d1.bias"])
    f2 = F.linear(f1, sd["depth_fc_d2.weight"], sd["depth_fc_d2.bias"])
    mid = F.batch_norm(f2, sd["depth_bn.running_mean"], sd["depth_bn.running_var"], sd["depth_bn.weight"], sd["depth_bn.bias"], False, 0.0, 1e-5)
    mid = F.leaky_relu(mid, 0.01)
    f3 = 0.5 * (F.linear(mid, sd["depth_fc_u2.weight"], sd["depth_fc_u2.bias"]) + f1)
    return 0.5 * (F.linear(f3, sd["depth_fc_u1.weight"], sd["depth_fc_u1.bias"]) + feat)


def joint_map_head(x_out, sd, bounds):
    # full_net.py:240-258, 376-379 (reg_joint_map): three conv3x3(+bias) + BN + ReLU on the trunk's map, a 1x1 conv to one map per
    # joint, then HeatmapIntegralJoint (lib/utils/integral.py:229-251): softmax over the positions, expected index / count, scaled
    # into the joint's [lower, upper]
    y = x_out
    for i in (0, 3, 6):
        y = F.conv2d(y, sd["joint_conv_layers.%d.weight" % i], sd["joint_conv_layers.%d.bias" % i], padding=1)
        p = "joint_conv_layers.%d." % (i + 1)
        y = F.relu(F.batch_norm(y, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5))
    y = F.conv2d(y, sd["joint_final_layer.weight"], sd["joint_final_layer.bias"])
    hm = F.softmax(y.reshape(y.shape[0], y.shape[1], -1), 2)
    hm = hm / hm.sum(dim=2, keepdim=True)
    n = hm.shape[-1]
    coord = (hm * torch.arange(n, dtype=torch.float32).reshape(1, 1, n)).sum(dim=2) / float(n)
    b = torch.as_tensor(bounds, dtype=torch.float32)
    return coord * (b[:, 1] - b[:, 0])[None] + b[:, 0][None]
