"""CPU restatement of the reference's evaluation tail. TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

Follows lib/utils/metrics.py of the reference: `batch_errors` <- compute_metrics_batch (metrics.py:8-118, after its FK call at
:29-34, whose result is an input here) and `summary` <- summary_add_pck (metrics.py:121-162). Pinned by
tests/golden/metrics_<robot>.npz, which holds what the reference's own two functions returned on the same inputs
(oracle/refrun/make_golden.py metrics)."""
import numpy as np

ADD_MM = (1, 5, 10, 20, 40, 60, 80, 100)                     # metrics.py:51 / 127
PCK_PX = (2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0)       # metrics.py:52 / 128


def project(K, xyz):
    """point_projection_from_3d (lib/utils/transforms.py:11-15): K @ p, divided by its last component."""
    h = np.einsum("bij,bkj->bki", K, xyz)
    return h[..., :2] / h[..., 2:3]


def batch_errors(pred_xyz, gt_xyz, gt_uv, K_original, gt_joint, pred_joint, root, robot_type):
    """float32 arrays in, the reference's 9-tuple out (same order as metrics.py:118)."""
    pred_xyz = np.asarray(pred_xyz, np.float32)
    gt_xyz = np.asarray(gt_xyz, np.float32)
    gt_uv = np.asarray(gt_uv, np.float32)
    pred_uv = project(np.asarray(K_original, np.float32), pred_xyz)                      # metrics.py:42
    e3 = np.sqrt(((pred_xyz - gt_xyz) ** 2).sum(-1))                                      # :55
    error3d = e3.mean(1)                                                                  # :57
    e2 = np.sqrt(((pred_uv - gt_uv) ** 2).sum(-1))                                        # :61
    inside = (gt_uv[..., 0] <= 640.0) & (gt_uv[..., 0] >= 0) & (gt_uv[..., 1] <= 480.0) & (gt_uv[..., 1] >= 0)   # :63
    with np.errstate(invalid="ignore", divide="ignore"):
        error2d = (e2 * inside).sum(1) / inside.sum(1)                                    # :64-67
        dis2d = (e2 * inside).sum(0) / inside.sum(0)                                      # :74-76
    dis3d = e3.mean(0)                                                                    # :73
    if pred_joint is not None:
        ej = np.abs(np.asarray(gt_joint, np.float32) - np.asarray(pred_joint, np.float32))   # :85
        l1_joint = ej.mean(0)                                                             # :86
        mean_joint = (ej[:, :-1] if robot_type == "panda" else ej).mean(1)                # :87-90
    else:
        l1_joint = np.zeros(np.asarray(gt_joint).shape[1], np.float32)                    # :92-93
        mean_joint = np.zeros(len(pred_xyz), np.float32)
    error_depth = np.abs(pred_xyz[:, root, 2] - gt_xyz[:, root, 2])                       # :97
    prel = pred_xyz[..., 2] - pred_xyz[:, root:root + 1, 2]                               # :100-101
    grel = gt_xyz[..., 2] - gt_xyz[:, root:root + 1, 2]
    rel = np.abs(prel - grel).mean(1)                                                     # :102-103
    p = pred_xyz.copy(); p[..., 2] = prel                                                 # :106-109
    g = gt_xyz.copy(); g[..., 2] = grel
    error3d_rel = np.sqrt(((p - g) ** 2).sum(-1)).mean(1)                                 # :110-112
    return error3d, error2d, dis3d, dis2d, l1_joint, mean_joint, error_depth, rel, error3d_rel


def _auc(values, stop, step):
    """Trapezoid of mean(values <= t) over t = arange(0, stop, step), divided by `stop` (metrics.py:131-140 / 143-152)."""
    t = np.arange(0.0, stop, step)
    frac = (values[None, :].astype(np.float64) <= t[:, None]).mean(1)
    return float(step * (frac.sum() - 0.5 * (frac[0] + frac[-1])) / stop)


def summary(dis3d, dis2d):
    """Dict with the reference's keys (metrics.py:154-162)."""
    d3 = np.asarray(dis3d, np.float32)
    d2 = np.asarray(dis2d, np.float32)
    out = {"ADD/mean": d3.mean(), "ADD/median": np.median(d3), "ADD/AUC": _auc(d3, 0.1, 0.00001),
           "ADD_2D/mean": d2.mean(), "ADD_2D/median": np.median(d2), "PCK/AUC": _auc(d2, 20.0, 0.01)}
    for mm in ADD_MM:
        out["ADD_%s_mm" % mm] = float((d3.astype(np.float64) <= mm * 1e-3).mean())
    for px in PCK_PX:
        out["PCK_%s_pixel" % px] = float((d2.astype(np.float64) <= px).mean())
    return out
