"""Evaluation tail on the GPU: drop-ins for compute_metrics_batch / summary_add_pck (lib/utils/metrics.py:8-118, 121-162).

The reference copies every prediction to the host and does this in numpy; here the batch errors are computed where the forward
left its outputs (libhrp_b200: hrp_metrics_batch, hrp_summary_add_pck), and a run accumulates its per-frame errors on the
device (ErrorLog) so that the only device->host traffic of an evaluation is the final summary. No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import capi
from .model import FkRobot, _ptr

ADD_MM = (1, 5, 10, 20, 40, 60, 80, 100)
PCK_PX = (2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0)
SUMMARY_KEYS = (["ADD/mean", "ADD/median", "ADD/AUC", "ADD_2D/mean", "ADD_2D/median", "PCK/AUC"] +
                ["ADD_%s_mm" % t for t in ADD_MM] + ["PCK_%s_pixel" % t for t in PCK_PX])


def _f32(t, device, shape):
    t = torch.as_tensor(t, dtype=torch.float32).to(device).contiguous()
    if tuple(t.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
    return t


def metrics_batch_device(robot, gt_keypoints3d, gt_keypoints2d, K_original, gt_joint, pred_joint=None, pred_rot=None,
                         pred_trans=None, pred_xyz_integral=None, reference_keypoint_id=None):
    """Device-resident form: returns (per_frame [B,6], dis3d [nkpt], dis2d [nkpt], l1_joint [dof]) CUDA tensors; per_frame
    columns = error3d, error2d, mean_jointerror, error_depth, batch_error_relative, error3d_relative. `robot`: an FkRobot
    (or anything with .fk_robot) for the same robot / root keypoint as the model."""
    fk = getattr(robot, "fk_robot", robot)
    if not isinstance(fk, FkRobot):
        raise TypeError("robot must be an hrp_b200 FkRobot (URDFRobot.get_keypoints_root runs on the device here)")
    prog = fk.program
    root = prog.root_kp if reference_keypoint_id is None else int(reference_keypoint_id)
    if root != prog.root_kp:
        raise ValueError("reference_keypoint_id=%d but the kinematic program is rooted at keypoint %d" % (root, prog.root_kp))
    dev = gt_keypoints3d.device if isinstance(gt_keypoints3d, torch.Tensor) and gt_keypoints3d.is_cuda else torch.device("cuda", torch.cuda.current_device())
    B = int(gt_keypoints3d.shape[0])
    nk, dof = prog.nkpt, prog.dof
    g3 = _f32(gt_keypoints3d, dev, (B, nk, 3))
    g2 = _f32(gt_keypoints2d, dev, (B, nk, 2))
    Ko = _f32(K_original, dev, (B, 3, 3))
    gq = _f32(gt_joint, dev, (B, dof))
    if pred_joint is None or pred_rot is None or pred_trans is None:          # metrics.py:22-26
        if pred_xyz_integral is None:
            raise ValueError("either pred_joint / pred_rot / pred_trans or pred_xyz_integral is needed")
        p3 = _f32(pred_xyz_integral, dev, (B, nk, 3))
        p2 = None                                                               # projected inside the kernels with K_original
        pq = None
    else:
        pq = _f32(pred_joint, dev, (B, dof))
        p3, p2 = fk.keypoints(pq, _f32(pred_rot, dev, (B, 6)), _f32(pred_trans, dev, (B, 3)), Ko)   # metrics.py:28-42
    per_frame = torch.empty(B, 6, device=dev, dtype=torch.float32)
    dis3d = torch.empty(nk, device=dev, dtype=torch.float32)
    dis2d = torch.empty(nk, device=dev, dtype=torch.float32)
    l1 = torch.empty(dof, device=dev, dtype=torch.float32)
    joint_cols = dof - 1 if fk.robot_type == "panda" else dof                   # metrics.py:87-90
    st = torch.cuda.current_stream(dev).cuda_stream
    null = C.c_void_p(0)
    capi.check(capi.lib().hrp_metrics_batch(_ptr(p3), _ptr(p2) if p2 is not None else null, _ptr(Ko), _ptr(pq) if pq is not None else null, _ptr(g3), _ptr(g2),
                                            _ptr(gq), B, nk, dof, root, joint_cols, _ptr(per_frame), _ptr(dis3d), _ptr(dis2d),
                                            _ptr(l1), C.c_void_p(st)))
    return per_frame, dis3d, dis2d, l1


def compute_metrics_batch(robot, gt_keypoints3d, gt_keypoints2d, K_original, gt_joint, **pred_kwargs):
    """Same call and the same 9 return values as the reference (metrics.py:8, :118): error3d [B], error2d [B] numpy arrays,
    dis3d / dis2d per keypoint, l1_jointerror per joint, mean_jointerror per frame (lists of floats), error_depth [B],
    batch_error_relative [B], error3d_relative [B]. One device->host copy of the packed result."""
    if pred_kwargs.get("pred_xy") is not None and pred_kwargs.get("pred_depth") is not None:        # metrics.py:15-18
        pred_kwargs = dict(pred_kwargs, pred_trans=torch.cat((pred_kwargs["pred_xy"], pred_kwargs["pred_depth"]), dim=-1))
    pf, d3, d2, l1 = metrics_batch_device(robot, gt_keypoints3d, gt_keypoints2d, K_original, gt_joint,
                                          pred_joint=pred_kwargs.get("pred_joint"), pred_rot=pred_kwargs.get("pred_rot"),
                                          pred_trans=pred_kwargs.get("pred_trans"), pred_xyz_integral=pred_kwargs.get("pred_xyz_integral"),
                                          reference_keypoint_id=pred_kwargs.get("reference_keypoint_id"))
    packed = torch.cat([pf.reshape(-1), d3, d2, l1]).cpu().numpy()
    B, nk, dof = pf.shape[0], d3.numel(), l1.numel()
    pf = packed[:B * 6].reshape(B, 6)
    o = B * 6
    dis3d, dis2d, l1j = packed[o:o + nk], packed[o + nk:o + 2 * nk], packed[o + 2 * nk:o + 2 * nk + dof]
    return (pf[:, 0].copy(), pf[:, 1].copy(), list(dis3d), dis2d.copy(), list(l1j), list(pf[:, 2]), pf[:, 3].copy(), pf[:, 4].copy(),
            pf[:, 5].copy())


class ErrorLog:
    """Per-frame error lists of a whole evaluation run, kept on the device (the reference extends Python lists batch by batch,
    scripts/test.py:214-225)."""

    def __init__(self, device=None, capacity=4096):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._buf = torch.empty(capacity, 6, device=self.device, dtype=torch.float32)
        self.n = 0

    def extend(self, per_frame):
        b = per_frame.shape[0]
        if self.n + b > self._buf.shape[0]:
            grown = torch.empty(max(2 * self._buf.shape[0], self.n + b), 6, device=self.device, dtype=torch.float32)
            grown[:self.n] = self._buf[:self.n]
            self._buf = grown
        self._buf[self.n:self.n + b] = per_frame
        self.n += b

    def column(self, i):
        return self._buf[:self.n, i].contiguous()

    def summary(self, relative=False):
        """summary_add_pck(alldis) (relative=False) or summary_add_pck(alldis_relative) (True: error3d_relative as dis3d)."""
        return summary_add_pck({"dis3d": self.column(5 if relative else 0), "dis2d": self.column(1)})


def summary_add_pck(alldis):
    """Drop-in for metrics.py:121-162. alldis['dis3d'] / ['dis2d']: CUDA tensors (kept there by ErrorLog), or anything
    torch.as_tensor accepts (lists of floats, as the reference accumulates them). Same keys as the reference."""
    def dev(v):
        t = v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v, np.float32))
        return t.to(device=torch.device("cuda", torch.cuda.current_device()) if not t.is_cuda else t.device, dtype=torch.float32).contiguous().reshape(-1)
    d3, d2 = dev(alldis["dis3d"]), dev(alldis["dis2d"])
    if d3.numel() != d2.numel():
        raise ValueError("dis3d and dis2d differ in length (%d, %d)" % (d3.numel(), d2.numel()))     # metrics.py:125
    n = d3.numel()
    L = capi.lib()
    ws = torch.empty(int(L.hrp_summary_workspace(n)), device=d3.device, dtype=torch.uint8)
    out = torch.empty(22, device=d3.device, dtype=torch.float64)
    st = torch.cuda.current_stream(d3.device).cuda_stream
    capi.check(L.hrp_summary_add_pck(_ptr(d3), _ptr(d2), n, _ptr(out), _ptr(ws), ws.numel(), C.c_void_p(st)))
    v = out.cpu().numpy()
    a, p = v[:11], v[11:]
    res = {"ADD/mean": np.float32(a[0]), "ADD/median": np.float32(a[1]), "ADD/AUC": float(a[2]),
           "ADD_2D/mean": np.float32(p[0]), "ADD_2D/median": np.float32(p[1]), "PCK/AUC": float(p[2])}
    for i, t in enumerate(ADD_MM):
        res["ADD_%s_mm" % t] = np.float64(a[3 + i])
    for i, t in enumerate(PCK_PX):
        res["PCK_%s_pixel" % t] = np.float64(p[3 + i])
    return res
