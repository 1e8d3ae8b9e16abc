"""Oracle: heatmap soft-argmax and the camera transforms around it (fp32, PyTorch CPU / numpy).

Follows lib/utils/integral.py:56-73,102-208 (softmax over D*H*W, marginals, expectation, fixroot),
lib/utils/transforms.py:17-21 (projection), 33-82 (uvd_to_xyz), 142-153 (uvz2xyz_singlepoint).
"""
import torch
import torch.nn.functional as F


def inverse_intrinsics(K):
    """integral.py:56-73: only fx, fy, cx, cy are used; computed in float64, stored fp32."""
    B = K.shape[0]
    inv = torch.zeros(B, 3, 3, dtype=torch.float32)
    fx, fy = K[:, 0, 0].double(), K[:, 1, 1].double()
    inv[:, 0, 0] = 1.0 / fx
    inv[:, 0, 2] = -K[:, 0, 2].double() / fx
    inv[:, 1, 1] = 1.0 / fy
    inv[:, 1, 2] = -K[:, 1, 2].double() / fy
    inv[:, 2, 2] = 1
    return inv


def soft_argmax_uvd(logits, nkpt, rootid, fixroot, depth_dim=64, hm=64, path="resnet50"):
    """[B, nkpt*D, H, W] -> uvd [B, nkpt, 3] in [-0.5, 0.5].

    path="resnet50": integral.py:116-151 (softmax, explicit re-normalisation by the sum, marginal * arange, sum).
    path="hrnet32":  integral.py:166-190 (softmax, NO re-normalisation, marginal.matmul(arange)). The two differ by the
    fp32 rounding of a 262 144-term softmax denominator (~1e-4 in uvd on peaked heatmaps)."""
    B = logits.shape[0]
    p = F.softmax(logits.reshape(B, nkpt, -1), 2)
    r = torch.arange(hm, dtype=torch.float32)
    if path == "resnet50":
        p = p / p.sum(2, keepdim=True)
        p = p.reshape(B, nkpt, depth_dim, hm, hm)
        x = (p.sum((2, 3)) * r).sum(2, keepdim=True)
        y = (p.sum((2, 4)) * r).sum(2, keepdim=True)
        z = (p.sum((3, 4)) * r).sum(2, keepdim=True)
    else:
        p = p.reshape(B, nkpt, depth_dim, hm, hm)
        x = p.sum((2, 3)).matmul(r.unsqueeze(-1))
        y = p.sum((2, 4)).matmul(r.unsqueeze(-1))
        z = p.sum((3, 4)).matmul(r.unsqueeze(-1))
    uvd = torch.cat((x / float(hm) - 0.5, y / float(hm) - 0.5, z / float(depth_dim) - 0.5), 2)
    if fixroot:
        uvd[:, rootid, 2] = 0.0
    return uvd


def uvd_to_xyz(uvd, K, root_z, image_size, depth_factor):
    """transforms.py:33-82 with return_relative=False. root_z: [B] absolute root depth (m)."""
    u = (uvd[:, :, 0] + 0.5) * image_size
    v = (uvd[:, :, 1] + 0.5) * image_size
    dz = uvd[:, :, 2] * depth_factor
    homo = torch.stack((u, v, torch.ones_like(u)), 2)
    ray = torch.matmul(inverse_intrinsics(K).unsqueeze(1), homo.unsqueeze(-1)).squeeze(3)
    return ray * (dz + root_z.reshape(-1, 1)).unsqueeze(-1)


def uvz_to_xyz(uv, z, K):
    """transforms.py:142-153."""
    v = torch.cat([uv * z, z], 1)
    return torch.matmul(inverse_intrinsics(K), v.unsqueeze(-1)).squeeze(-1)


def project(K, xyz):
    """transforms.py:17-21 (vectorised): (K @ p)[:2] / (K @ p)[2]."""
    h = torch.matmul(K.unsqueeze(1), xyz.unsqueeze(-1)).squeeze(-1)
    return h[..., :2] / h[..., 2:3]
