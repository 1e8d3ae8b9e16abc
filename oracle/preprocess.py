"""Oracle: the reference's CPU input preparation for one inference crop. TEST INFRASTRUCTURE ONLY.

Restates, with the same numpy / torch-CPU calls the reference makes:
  resize_image                       lib/dataset/roboutils.py:142-171   (box pasted into a zero square, K shifted)
  CropResizeToAspectAugmentation     lib/dataset/augmentations.py:189-262 (bilinear on /255 floats, back to uint8)
  get_K_crop_resize                  lib/utils/geometries.py:360-402
  bbox_transform + clipping          lib/dataset/roboutils.py:248-263, lib/dataset/dream.py:445-449
  k_value                            lib/core/function.py:98-110 (= scripts/test.py:143-153)
Pinned against the reference's own functions by tests/golden/preprocess.npz (oracle/refrun/make_golden.py preprocess()).
"""
import numpy as np
import torch
import torch.nn.functional as F


def get_K_crop_resize(K, boxes, crop_resize):
    K = K.float()
    boxes = boxes.float()
    new_K = K.clone()
    crop_resize = torch.tensor(crop_resize, dtype=torch.float)
    final_width, final_height = max(crop_resize), min(crop_resize)
    crop_width = boxes[:, 2] - boxes[:, 0]
    crop_height = boxes[:, 3] - boxes[:, 1]
    crop_cj = (boxes[:, 0] + boxes[:, 2]) / 2
    crop_ci = (boxes[:, 1] + boxes[:, 3]) / 2
    cx = K[:, 0, 2] + (crop_width - 1) / 2 - crop_cj
    cy = K[:, 1, 2] + (crop_height - 1) / 2 - crop_ci
    center_x = (crop_width - 1) / 2
    center_y = (crop_height - 1) / 2
    scale_x = final_width / crop_width
    scale_y = final_height / crop_height
    new_K[:, 0, 0] = scale_x * K[:, 0, 0]
    new_K[:, 1, 1] = scale_y * K[:, 1, 1]
    new_K[:, 0, 2] = (final_width - 1) / 2 + scale_x * (cx - center_x)
    new_K[:, 1, 2] = (final_height - 1) / 2 + scale_y * (cy - center_y)
    return new_K


def crop_resize_one(frame, crop_box, K, k_box=None, size=256):
    """frame [H,W,3] uint8, crop_box (wmin,hmin,wmax,hmax) ints, K [3,3] float64 -> (crop uint8 [3,size,size], K' fp32 [3,3],
    k_value fp32 scalar or None)."""
    wmin, hmin, wmax, hmax = (int(v) for v in crop_box)
    S = int(max(wmax - wmin, hmax - hmin))
    square = np.zeros((S, S, 3), np.uint8)
    x_off = int((S - (wmax - wmin)) // 2)
    y_off = int((S - (hmax - hmin)) // 2)
    square[y_off:y_off + (hmax - hmin), x_off:x_off + (wmax - wmin)] = frame[hmin:hmax, wmin:wmax]
    K = np.array(K, np.float64)
    K_orig = K.copy()
    K[0, 2] -= (wmin - x_off)
    K[1, 2] -= (hmin - y_off)
    images = (torch.as_tensor(square).float() / 255).unsqueeze(0).permute(0, 3, 1, 2)
    if (S, S) != (size, size):
        x0, y0 = S / 2, S / 2
        box = torch.tensor([x0 - S / 2, y0 - S / 2, x0 + S / 2, y0 + S / 2])
        images = F.interpolate(images, size=(size, size), mode="bilinear", align_corners=False)
        Kn = get_K_crop_resize(torch.tensor(K).unsqueeze(0), box.unsqueeze(0), (size, size))[0].numpy()
    else:                                       # augmentations.py:193-195: already at the target size, nothing changes
        Kn = K.astype(np.float32)
    crop = (images[0] * 255).to(torch.uint8).numpy()
    kv = None
    if k_box is not None:
        bx0, by0, bx1, by1 = (float(v) for v in k_box)
        corners = np.array([[bx0, by0, 1.0], [bx1, by0, 1.0], [bx1, by1, 1.0], [bx0, by1, 1.0]])
        new = np.matmul(Kn, np.matmul(np.linalg.inv(K_orig), corners.T)).T
        t = np.array([np.clip(new[0, 0], 0, size), np.clip(new[0, 1], 0, size), np.clip(new[1, 0], 0, size), np.clip(new[2, 1], 0, size)])
        t = torch.FloatTensor(np.array([max(0, t[0]), max(0, t[1]), min(size, t[2]), min(size, t[3])]))
        fx, fy = torch.tensor(Kn[0, 0]), torch.tensor(Kn[1, 1])
        area = torch.max(torch.abs(t[2] - t[0]), torch.abs(t[3] - t[1])) ** 2
        kv = torch.sqrt(fx * fy * torch.tensor(1000.0) * torch.tensor(1000.0) / area).numpy().astype(np.float32)
    return crop, Kn.astype(np.float32), kv
