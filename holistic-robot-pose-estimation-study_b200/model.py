"""Host-side mirror of the reference model interface, backed by the CUDA library.

`HoliRobPoseB200` is a drop-in for `RootNetwithRegInt` (lib/models/full_net.py:17-466) on the inference path:
same constructor intent (robot type + the ctor-relevant config keys), `load_state_dict` with the reference's key names
(optionally `module.`-prefixed, scripts/fullnet_test.py:193-198), `forward(x_reg, x_root, k_value, K)` returning the same
8-tuple in the same order / shapes / dtype (fp32 on the inputs' CUDA device), plus the `forward(images, K)` dict
convenience of BASELINE.json's north_star. PyTorch is used for device memory and streams only.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import arch, capi, consts, urdf

UNSUPPORTED = {  # ctor switches of the reference that no shipped config enables and this path does not build -> explicit error
    "use_rpmg": False,
}


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class FkRobot:
    """Drop-in for URDFRobot.get_keypoints / get_keypoints_root (+ projection) on CUDA tensors."""

    def __init__(self, robot_type, urdf_text=None):
        self.robot_type = robot_type
        self.parsed, self.program = urdf.load_robot(robot_type, urdf_text)
        self._struct, self._keep = capi.fk_program_struct(self.program)
        h = C.c_void_p()
        capi.check(capi.lib().hrp_fk_create(C.byref(self._struct), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().hrp_fk_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def keypoints(self, q, rot6d, trans, K):
        """q [N,dof], rot6d [N,6], trans [N,3], K [N,3,3] CUDA fp32 -> (xyz [N,nkpt,3], uv [N,nkpt,2])."""
        n = q.shape[0]
        for t, w in ((q, self.program.dof), (rot6d, 6), (trans, 3)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape == (n, w)):
                raise ValueError("FkRobot.keypoints: expected CUDA fp32 [N,%d], got %s %s" % (w, t.dtype, tuple(t.shape)))
        if not (K.is_cuda and K.dtype == torch.float32 and K.shape == (n, 3, 3)):
            raise ValueError("FkRobot.keypoints: K must be CUDA fp32 [N,3,3]")
        q, rot6d, trans, K = (t.contiguous() for t in (q, rot6d, trans, K))
        xyz = torch.empty(n, self.program.nkpt, 3, device=q.device, dtype=torch.float32)
        uv = torch.empty(n, self.program.nkpt, 2, device=q.device, dtype=torch.float32)
        st = torch.cuda.current_stream(q.device).cuda_stream
        capi.check(capi.lib().hrp_fk_project(self._h, _ptr(q), _ptr(rot6d), _ptr(trans), _ptr(K), n, _ptr(xyz), _ptr(uv),
                                             C.c_void_p(st)))
        return xyz, uv


def soft_argmax(heatmap, nkpt, K=None, root_z=None, depth_factor=1.3, image_size=256.0, rootid=0, fixroot=True):
    """Drop-in for HeatmapIntegralPose.forward (lib/utils/integral.py:102-208).

    heatmap [B, nkpt*D, H, W] CUDA fp32 (reference layout) -> (uvd [B,nkpt,3], xyz [B,nkpt,3] or None)."""
    if not (heatmap.is_cuda and heatmap.dtype == torch.float32 and heatmap.dim() == 4):
        raise ValueError("soft_argmax: heatmap must be CUDA fp32 [B, nkpt*D, H, W]")
    B, CD, H, W = heatmap.shape
    if nkpt <= 0 or CD % nkpt:
        raise ValueError("soft_argmax: channel count %d is not a multiple of nkpt=%d" % (CD, nkpt))
    D = CD // nkpt
    heatmap = heatmap.contiguous()
    uvd = torch.empty(B, nkpt, 3, device=heatmap.device, dtype=torch.float32)
    xyz = None
    kp = rp = C.c_void_p(0)
    if K is not None:
        K = K.contiguous().float()
        root_z = root_z.contiguous().float().reshape(-1)
        xyz = torch.empty(B, nkpt, 3, device=heatmap.device, dtype=torch.float32)
        kp, rp = _ptr(K), _ptr(root_z)
    L = capi.lib()
    nbytes = L.hrp_softargmax3d_workspace(B, nkpt, D, H, W)
    ws = torch.empty(max(nbytes, 16), device=heatmap.device, dtype=torch.uint8)
    st = torch.cuda.current_stream(heatmap.device).cuda_stream
    capi.check(L.hrp_softargmax3d(_ptr(heatmap), B, nkpt, D, H, W, kp, rp, depth_factor, image_size, rootid,
                                  int(bool(fixroot)), _ptr(uvd), _ptr(xyz) if xyz is not None else C.c_void_p(0),
                                  _ptr(ws), nbytes, C.c_void_p(st)))
    return uvd, xyz


def crop_resize(frames, crop_box, K, k_box=None):
    """Input side of the boundary on the device (the reference does this on the CPU in its DataLoader:
    lib/dataset/dream.py:415-449, roboutils.py:142-171,248-263, augmentations.py:189-262, geometries.py:360-402,
    lib/core/function.py:98-110).

    frames [B,Hf,Wf,3] uint8 CUDA (HWC camera images), crop_box [B,4] int32 (wmin,hmin,wmax,hmax), K [B,3,3] fp32,
    k_box [B,4] fp32 strict robot box in frame coordinates (optional)
    -> (crops [B,3,256,256] uint8 -- feed them to forward_dict / HostPipeline, which apply the `/255.` --, K' [B,3,3],
        k_value [B] or None)."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3):
        raise ValueError("crop_resize: frames must be a CUDA uint8 tensor [B,H,W,3]")
    B, Hf, Wf, _ = frames.shape
    dev = frames.device
    frames = frames.contiguous()
    crop_box = torch.as_tensor(crop_box, device=dev).to(torch.int32).reshape(B, 4).contiguous()
    K = torch.as_tensor(K, device=dev).float().reshape(B, 3, 3).contiguous()
    crops = torch.empty(B, 3, 256, 256, device=dev, dtype=torch.uint8)
    K_out = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
    kv, kb = None, C.c_void_p(0)
    kvp = C.c_void_p(0)
    if k_box is not None:
        k_box = torch.as_tensor(k_box, device=dev).float().reshape(B, 4).contiguous()
        kv = torch.empty(B, device=dev, dtype=torch.float32)
        kb, kvp = _ptr(k_box), _ptr(kv)
    st = torch.cuda.current_stream(dev).cuda_stream
    capi.check(capi.lib().hrp_crop_resize_u8(_ptr(frames), B, Hf, Wf, _ptr(crop_box), kb, _ptr(K), _ptr(crops), _ptr(K_out), kvp,
                                              C.c_void_p(st)))
    return crops, K_out, kv


def conv2d_nhwc(x, weight, bias=None, residual=None, stride=1, pad=0, relu=False, precision="fp32"):
    """Single conv through the library (layer-level parity tests). x NHWC, weight OIHW, CUDA fp32."""
    B, Hi, Wi, Cin = x.shape
    Cout, _, KH, KW = weight.shape
    Ho, Wo = (Hi + 2 * pad - KH) // stride + 1, (Wi + 2 * pad - KW) // stride + 1
    out = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.float32)
    st = torch.cuda.current_stream(x.device).cuda_stream
    z = C.c_void_p(0)
    capi.check(capi.lib().hrp_conv2d_nhwc(_ptr(x.contiguous()), _ptr(weight.contiguous()),
                                          _ptr(bias.contiguous()) if bias is not None else z,
                                          _ptr(residual.contiguous()) if residual is not None else z, _ptr(out),
                                          B, Hi, Wi, Cin, Cout, KH, KW, stride, pad, int(relu), capi.PREC[precision],
                                          C.c_void_p(st)))
    return out


def basic_block_nhwc(x, w1, b1, w2, b2):
    """One HRNet BasicBlock through the fused bf16 kernel (layer-level parity tests). x NHWC, weights OIHW, CUDA fp32."""
    nb, hh, ww, ch = x.shape
    out = torch.empty_like(x)
    st = torch.cuda.current_stream(x.device).cuda_stream
    z = C.c_void_p(0)
    capi.check(capi.lib().hrp_basic_block_nhwc(_ptr(x.contiguous()), _ptr(w1.contiguous()), _ptr(b1.contiguous()) if b1 is not None else z,
                                               _ptr(w2.contiguous()), _ptr(b2.contiguous()) if b2 is not None else z, _ptr(out),
                                               nb, hh, ww, ch, C.c_void_p(st)))
    return out


def basic_chain_nhwc(x, weights, biases):
    """A chain of HRNet BasicBlocks through the one-launch branch kernel (layer-level parity tests). x NHWC fp32;
    weights [2*nblocks, C, C, 3, 3], biases [2*nblocks, C] or None, CUDA fp32."""
    nb, hh, ww, ch = x.shape
    out = torch.empty_like(x)
    st = torch.cuda.current_stream(x.device).cuda_stream
    z = C.c_void_p(0)
    capi.check(capi.lib().hrp_basic_chain_nhwc(_ptr(x.contiguous()), _ptr(weights.contiguous()),
                                               _ptr(biases.contiguous()) if biases is not None else z, weights.shape[0] // 2,
                                               _ptr(out), nb, hh, ww, ch, C.c_void_p(st)))
    return out


class HoliRobPoseB200(torch.nn.Module):
    """CUDA drop-in for RootNetwithRegInt (inference forward only)."""

    def __init__(self, robot_type, cfg=None, device=None, precision="fp32", **kw):
        super().__init__()
        cfg = dict(cfg or {}, **kw)
        if robot_type not in consts.ROBOTS:
            raise ValueError("Robot type %s is not supported." % robot_type)       # full_net.py:55
        for k, v in UNSUPPORTED.items():
            if cfg.get(k, v) != v:
                raise NotImplementedError("config %s=%r is not supported by the B200 path (shipped value: %r)" % (k, cfg[k], v))
        if int(cfg.get("rotation_dim", 6)) != 6:
            raise NotImplementedError("rotation_dim=%r: only the 6-D representation is supported" % cfg.get("rotation_dim"))
        spec = consts.ROBOTS[robot_type]
        self.robot_type = robot_type
        self.backbone_name = cfg.get("backbone_name", "resnet50")
        self.rootnet_backbone_name = cfg.get("rootnet_backbone_name", "hrnet32")
        if self.backbone_name not in capi.BACKBONE:
            raise NotImplementedError("backbone_name=%r (supported: resnet50, hrnet32)" % self.backbone_name)
        if self.rootnet_backbone_name not in ("hrnet", "hrnet32"):
            raise NotImplementedError("rootnet_backbone_name=%r (supported: hrnet32)" % self.rootnet_backbone_name)
        if precision not in capi.PREC:
            raise ValueError("precision must be one of %s" % list(capi.PREC))
        self.precision = precision
        self.n_iter = int(cfg.get("n_iter", consts.N_ITER))
        self.image_size = float(cfg.get("other_image_size", consts.IMAGE_SIZE))
        self.bbox_3d_shape = tuple(cfg.get("bbox_3d_shape", spec["bbox_3d"]))
        self.reference_keypoint_id = int(cfg.get("reference_keypoint_id", spec["ref_kp"]))
        self.fix_root = bool(cfg.get("fix_root", True))
        # constructor variants outside the shipped configuration (full_net.py:107-131, 149-164)
        self.direct_reg_rot = bool(cfg.get("direct_reg_rot", False))
        self.rot_iterative_matmul = bool(cfg.get("rot_iterative_matmul", False))
        self.add_fc = bool(cfg.get("add_fc", False))
        self.multi_kp = bool(cfg.get("multi_kp", False))
        self.kps_need_depth = list(cfg.get("kps_need_depth") or []) if self.multi_kp else [self.reference_keypoint_id]   # full_net.py:150-151
        if self.multi_kp and self.reference_keypoint_id not in self.kps_need_depth:
            raise ValueError("%d is not in list" % self.reference_keypoint_id)      # what kps_need_depth.index raises, full_net.py:328
        self.depth_num = len(self.kps_need_depth)
        self.reg_joint_map = bool(cfg.get("reg_joint_map", False))
        self.joint_conv_dim = [int(v) for v in (cfg.get("joint_conv_dim") or [])] if self.reg_joint_map else []
        if self.reg_joint_map:
            if self.backbone_name not in ("resnet", "resnet50"):
                raise NotImplementedError("reg_joint_map reads the ResNet trunk's feature map (full_net.py:377); backbone_name=%r has none" % self.backbone_name)
            if len(self.joint_conv_dim) != 3 or any(d <= 0 or d % 32 for d in self.joint_conv_dim):
                raise NotImplementedError("joint_conv_dim=%r: three positive multiples of 32 are supported" % (self.joint_conv_dim,))
        self.dof, self.nkpt = spec["dof"], spec["nkpt"]
        self.num_joints = self.nkpt
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda", 0)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HoliRobPoseB200 runs on CUDA devices only (got %s); there is no CPU fallback" % self.device)
        if self.device.index is None:                                    # 'cuda' = the process's CURRENT device, not GPU 0
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        parsed = urdf.Robot(open(consts.urdf_path(robot_type)).read())
        self.program = urdf.compile_program(parsed, urdf.keypoint_frames(robot_type, parsed), self.reference_keypoint_id,
                                            spec["joints"])
        self._prog_struct, self._prog_keep = capi.fk_program_struct(self.program)
        depth_factor = float(np.float32(self.bbox_3d_shape[2]) * np.float32(1e-3))       # integral.py:96-97
        self._cfg = capi.Config(capi.BACKBONE[self.backbone_name], capi.PREC[precision], self.n_iter, int(self.fix_root),
                                self.image_size, depth_factor, int(self.direct_reg_rot), int(self.rot_iterative_matmul),
                                int(self.add_fc), self.depth_num if self.multi_kp else 0,
                                self.kps_need_depth.index(self.reference_keypoint_id) if self.multi_kp else 0,
                                int(self.reg_joint_map), (C.c_int32 * 3)(*(self.joint_conv_dim or [0, 0, 0])),
                                (C.c_float * 32)(*[float(v) for lo_hi in spec["bounds"] for v in lo_hi]))   # const.py:239-284
        h = C.c_void_p()
        capi.check(capi.lib().hrp_create(C.byref(self._cfg), C.byref(self._prog_struct), self.device.index, C.byref(h)))
        self._h = h
        self._finalized = False
        self._out = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                capi.lib().hrp_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    # ---- weights ----------------------------------------------------------------------------------------------------
    def expected_tensors(self):
        L = capi.lib()
        out = []
        shape = (C.c_int64 * 4)()
        nd = C.c_int()
        for i in range(L.hrp_num_weights(self._h)):
            capi.check(L.hrp_weight_shape(self._h, i, shape, C.byref(nd)))
            out.append((L.hrp_weight_name(self._h, i).decode(), tuple(shape[k] for k in range(nd.value))))
        return out

    def load_state_dict(self, state_dict, strict=True):
        """Accepts a reference state dict (tensors or numpy arrays; optional `module.` prefix)."""
        if self._finalized:
            raise RuntimeError("weights are already loaded; build a new HoliRobPoseB200 to load another checkpoint")
        L = capi.lib()
        expected = dict(self.expected_tensors())
        seen = self.__dict__.setdefault("_seen", set())     # accumulates over strict=False loads (pretrained DepthNet, then the rest)
        unexpected = []
        for k, v in state_dict.items():
            if k.startswith("module."):
                k = k[len("module."):]
            if k not in expected:
                unexpected.append(k)
                continue
            if k.endswith("num_batches_tracked"):
                seen.add(k)
                continue
            a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            capi.check(L.hrp_set_weight(self._h, k.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim, 0))
            seen.add(k)
        missing = [k for k in expected if k not in seen and not k.endswith("num_batches_tracked")]
        if strict and (missing or unexpected):
            raise RuntimeError("load_state_dict: missing %s, unexpected %s" % (missing[:5], unexpected[:5]))
        if missing:
            # strict=False on a torch module leaves absent tensors at their constructor values; this handle has no
            # constructor values (weights only ever come from a checkpoint), so it stays open for the next load
            self._pending_missing = missing
            return missing, unexpected
        capi.check(L.hrp_finalize_weights(self._h))
        self._finalized = True
        self._pending_missing = []
        return missing, unexpected

    @staticmethod
    def read_checkpoint(path_or_dict, map_location="cpu"):
        """Reference checkpoint file -> plain state dict. Accepts what `save_checkpoint` writes
        (lib/utils/utils.py:248-254: {'epoch', 'auc_add', 'model_state_dict', 'optimizer_state_dict',
        'lr_scheduler_last_epoch'}), a bare state dict, DataParallel / DDP `module.` prefixes
        (scripts/fullnet_test.py:193-198)."""
        ck = path_or_dict
        if not isinstance(ck, dict):
            ck = torch.load(path_or_dict, map_location=map_location, weights_only=False)
        sd = ck["model_state_dict"] if "model_state_dict" in ck else ck
        return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}

    def load_checkpoint(self, path_or_dict, pretrained_rootnet=None, strict=True):
        """`torch.load(path)['model_state_dict']` + `load_state_dict` as the reference's evaluators do
        (scripts/test.py, scripts/fullnet_test.py:186-198, lib/utils/utils.py:198-200).

        pretrained_rootnet: optional DepthNet pre-training checkpoint (lib/models/depth_net.py `RootNet`), merged the way
        the factory does (full_net.py:486-500): keys `backbone.*` are re-keyed to `rootnet_backbone.*`, everything is
        loaded with strict=False semantics, i.e. only names this network has are taken (`depth_layer.*` among them) and
        they override the same names of the main checkpoint."""
        sd = dict(self.read_checkpoint(path_or_dict))
        if pretrained_rootnet is not None:
            expected = dict(self.expected_tensors())
            for k, v in self.read_checkpoint(pretrained_rootnet).items():
                nk = k.replace("backbone", "rootnet_backbone") if k.startswith("backbone") else k   # full_net.py:494-498
                if nk in expected:
                    sd[nk] = v
        return self.load_state_dict(sd, strict=strict)

    # ---- forward -----------------------------------------------------------------------------------------------------
    def _record(self, B, device):
        key = (B, device)
        if key not in self._out:
            offs = (C.c_int64 * (capi.NUM_FIELDS + 1))()
            capi.check(capi.lib().hrp_output_offsets(self._h, B, offs))
            self._out[key] = list(offs)
        return self._out[key]

    def forward_record(self, x_reg, x_root, k_value, K, init_pose=None, init_rot=None, times=None):
        """Runs the network; returns (flat fp32 record tensor, field offsets). init_pose [B,dof] / init_rot [B,6]:
        optional initial states of the refinement heads (full_net.py:268-272). times: a list that receives
        (time_root, time_other, time_whole) in seconds -- the reference's test_fps mode, which synchronises."""
        if not self._finalized:
            raise RuntimeError("forward before load_state_dict" + (
                " (the last load left tensors missing: %s ...)" % self._pending_missing[:3] if getattr(self, "_pending_missing", None) else ""))
        B = x_reg.shape[0]
        for t, shp in ((x_reg, (B, 3, 256, 256)), (x_root, (B, 3, 256, 256)), (K, (B, 3, 3))):
            if not t.is_cuda or tuple(t.shape) != shp:
                raise ValueError("expected a CUDA tensor of shape %s, got %s on %s" % (shp, tuple(t.shape), t.device))
            if t.device != self.device:
                raise ValueError("input lives on %s but this model was built for %s" % (t.device, self.device))
        if B == 0:                                                           # torch modules pass empty batches through
            if times is not None:
                times.append((0.0, 0.0, 0.0))
            return torch.empty(0, device=x_reg.device, dtype=torch.float32), [0] * (capi.NUM_FIELDS + 1)
        same = x_root is x_reg                                               # views that merely START at the same address differ
        u8 = x_reg.dtype == torch.uint8 and x_root.dtype == torch.uint8 and init_pose is None and init_rot is None and times is None
        if u8:           # the DataLoader's uint8 crops: `/ 255.` (scripts/test.py:93-96) happens inside the stem's input pack
            x_reg = x_reg.contiguous()
            x_root = x_reg if same else x_root.contiguous()
        else:
            if x_reg.dtype == torch.uint8 or x_root.dtype == torch.uint8:
                raise ValueError("uint8 images are taken by forward_dict / forward_record / HostPipeline without init_pose / init_rot / "
                                 "test_fps; the reference-signature call expects float images already divided by 255 (scripts/test.py:93-96)")
            x_reg = x_reg.float().contiguous()                               # full_net.py:265-266
            x_root = x_reg if same else x_root.float().contiguous()
        k_value = torch.as_tensor(k_value, device=x_reg.device).float().reshape(B).contiguous()
        K = K.float().contiguous()
        null = C.c_void_p(0)
        ip = ir = null
        if init_pose is not None:
            init_pose = torch.as_tensor(init_pose, device=x_reg.device).float().expand(B, self.dof).contiguous()
            ip = _ptr(init_pose)
        if init_rot is not None:
            init_rot = torch.as_tensor(init_rot, device=x_reg.device).float().expand(B, 6).contiguous()
            ir = _ptr(init_rot)
        offs = self._record(B, x_reg.device)
        rec = torch.empty(offs[-1], device=x_reg.device, dtype=torch.float32)
        st = torch.cuda.current_stream(x_reg.device).cuda_stream
        L = capi.lib()
        if u8:
            capi.check(L.hrp_forward_u8(self._h, _ptr(x_reg), _ptr(x_root), _ptr(k_value), _ptr(K), B, _ptr(rec), C.c_void_p(st)))
        elif times is not None:
            ms = (C.c_float * 3)()
            capi.check(L.hrp_forward_timed(self._h, _ptr(x_reg), _ptr(x_root), _ptr(k_value), _ptr(K), ip, ir, B, _ptr(rec), ms,
                                           C.c_void_p(st)))
            times.append(tuple(1e-3 * v for v in ms))
        elif init_pose is not None or init_rot is not None:
            capi.check(L.hrp_forward_ex(self._h, _ptr(x_reg), _ptr(x_root), _ptr(k_value), _ptr(K), ip, ir, B, _ptr(rec),
                                        C.c_void_p(st)))
        else:
            capi.check(L.hrp_forward(self._h, _ptr(x_reg), _ptr(x_root), _ptr(k_value), _ptr(K), B, _ptr(rec), C.c_void_p(st)))
        return rec, offs

    def _fields(self, rec, offs, B):
        w = (self.dof, 6, 3, 2, 1, self.nkpt * 3, self.nkpt * 3, self.nkpt * 3, self.nkpt * 2, self.nkpt * 2,
             self.depth_num if self.multi_kp else 0)
        out = []
        for f in range(capi.NUM_FIELDS):
            t = rec[offs[f]:offs[f] + B * w[f]].view(B, w[f])
            if f in (5, 6, 7):
                t = t.view(B, self.nkpt, 3)
            elif f in (8, 9):
                t = t.view(B, self.nkpt, 2)
            out.append(t)
        return out

    def forward(self, x_reg_input, x_root_input=None, k_value=None, K=None, init_pose=None, init_rot=None,
                test_fps=False):
        """Reference signature -> 8-tuple (full_net.py:262, 466), or 9 values with test_fps=True: the 8 tensors and
        (time_root, time_other, time_whole) in seconds (full_net.py:459-460, unpacked by scripts/test.py:161-162).
        `model(images, K)` / `model(images, K=K)` -> dict."""
        if K is None and (k_value is None) and x_root_input is not None and x_root_input.dim() == 3:
            return self.forward_dict(x_reg_input, x_root_input)              # model(images, K)
        if x_root_input is None:
            return self.forward_dict(x_reg_input, K, k_value)
        B = x_reg_input.shape[0]
        times = [] if test_fps else None
        rec, offs = self.forward_record(x_reg_input, x_root_input, k_value, K, init_pose, init_rot, times)
        f = self._fields(rec, offs, B)
        if self.multi_kp and not test_fps:                                   # full_net.py:462-464: pred_depths rides along
            return tuple(f[:5]) + (f[10],) + tuple(f[5:8])
        return tuple(f[:8]) + ((times[0],) if test_fps else ())

    def forward_dict(self, images, K, k_value=None):
        """north_star convenience: dict of 2-D/3-D keypoints, joint angles, root depth and camera-frame pose.
        images: float [B,3,256,256] in [0,1], or the DataLoader's uint8 crops (then `/ 255.` happens on the device)."""
        B = images.shape[0]
        if k_value is None:                                                  # scripts/real_test.py:285-289, full-frame bbox
            k_value = torch.sqrt(K[:, 0, 0] * K[:, 1, 1] * 1000.0 * 1000.0 / (self.image_size * self.image_size))
        rec, offs = self.forward_record(images, images, k_value, K)
        f = self._fields(rec, offs, B)
        out = dict(zip(capi.FIELD_NAMES, f))
        if self.multi_kp:
            out["depths"] = f[10]
        return out

    def launch_count(self):
        return int(capi.lib().hrp_launch_count(self._h))

    def set_option(self, name, value):
        capi.check(capi.lib().hrp_set_option(self._h, name.encode(), int(value)))

    def release_plans(self):
        """Free every cached per-batch-size plan (workspace + CUDA graph); the next forward re-plans."""
        capi.check(capi.lib().hrp_release_plans(self._h))

    def debug_tensor(self, name, B):
        n = C.c_int64()
        L = capi.lib()
        capi.check(L.hrp_debug_tensor(self._h, name.encode(), B, C.c_void_p(0), C.byref(n), C.c_void_p(0)))
        t = torch.empty(n.value, device=self.device, dtype=torch.float32)
        st = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(L.hrp_debug_tensor(self._h, name.encode(), B, _ptr(t), C.byref(n), C.c_void_p(st)))
        return t

    def profile(self, x_reg, x_root, k_value, K):
        """Per-kernel-class device time of one un-graphed forward (CUDA events around every launch)."""
        B = x_reg.shape[0]
        offs = self._record(B, x_reg.device)
        rec = torch.empty(offs[-1], device=x_reg.device, dtype=torch.float32)
        ms = (C.c_float * capi.NUM_CLASSES)()
        ln = (C.c_int64 * capi.NUM_CLASSES)()
        fl = (C.c_double * capi.NUM_CLASSES)()
        st = torch.cuda.current_stream(x_reg.device).cuda_stream
        k_value = torch.as_tensor(k_value, device=x_reg.device).float().reshape(B).contiguous()
        capi.check(capi.lib().hrp_forward_profile(self._h, _ptr(x_reg.contiguous()), _ptr(x_root.contiguous()), _ptr(k_value),
                                                  _ptr(K.contiguous()), B, _ptr(rec), ms, ln, fl, C.c_void_p(st)))
        return {capi.CLASS_NAMES[i]: dict(ms=ms[i], launches=ln[i], flops=fl[i]) for i in range(capi.NUM_CLASSES)}


class HostPipeline:
    """Streaming inference from HOST batches: `submit()` enqueues the host->device copy of one batch on a copy stream
    (double-buffered device staging), the forward on the compute stream behind it and the device->host copy of the packed
    output record; `result()` blocks on that batch only. With `depth` slots (default 4, the library's number of plans
    per batch size) uploads overlap forwards and the low-parallelism tail of batch i overlaps the head of batch i+1
    (50 MB of fp32 images per 64 frames is ~0.9 ms over PCIe 5, 15 % of the forward at batch 64). Inputs should be
    pinned (`torch.Tensor.pin_memory()`), otherwise the copies serialise with the host.

        pipe = HostPipeline(model, batch=64)
        t = pipe.submit(images, K, k_value)          # images [B,3,256,256] fp32 in [0,1] or uint8 crops (x_reg == x_root, real_test.py:282)
        out = pipe.result(t)                         # dict of pinned-host views, valid until the slot is reused
    """

    def __init__(self, model, batch, depth=4, post=None):
        self.model, self.B, self.depth, self.post = model, int(batch), int(depth), post
        dev = model.device
        self.copy_stream = torch.cuda.Stream(dev)
        # one compute stream per slot: the library keeps two plans (workspace + graph) per batch size and uses them
        # round-robin, so forwards enqueued on different streams overlap (the tail of batch i with the head of batch i+1)
        self.compute = [torch.cuda.Stream(dev) for _ in range(depth)]
        # device staging: TWO input sets per slot, used alternately, so the upload of batch n never waits for the forward of
        # batch n - depth (which still reads the slot's other set while 'depth' forwards are in flight); allocated on first
        # use in the dtype the caller submits (fp32 or uint8)
        self.nstage = 2 * depth
        self.img = [None] * self.nstage
        self.K = [torch.empty(self.B, 3, 3, device=dev) for _ in range(self.nstage)]
        self.kv = [torch.empty(self.B, device=dev) for _ in range(self.nstage)]
        self.offs = model._record(self.B, dev)
        # pinned result buffers up front: a cudaHostAlloc inside the stream of submits synchronises the whole device
        self.host = [torch.empty(self.offs[-1], dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(self.nstage)]
        self.ev_free = [None] * self.nstage    # forward of the batch that last used this input set has finished
        self.ev_done = [None] * depth
        self.n = 0

    def submit(self, images, K, k_value=None):
        s = self.n % self.depth               # compute stream, plan and result buffer
        g = self.n % self.nstage              # input staging set
        self.n += 1
        if k_value is None:
            k_value = torch.sqrt(K[:, 0, 0] * K[:, 1, 1] * 1000.0 * 1000.0 / (self.model.image_size * self.model.image_size))
        compute = self.compute[s]
        compute.wait_stream(torch.cuda.current_stream(self.model.device))
        with torch.cuda.stream(self.copy_stream):
            if self.ev_free[g] is not None:
                self.copy_stream.wait_event(self.ev_free[g])
            if self.img[g] is None or self.img[g].dtype != images.dtype:
                self.img[g] = torch.empty(self.B, 3, 256, 256, device=self.model.device, dtype=images.dtype)
            self.img[g].copy_(images, non_blocking=True)
            self.K[g].copy_(K, non_blocking=True)
            self.kv[g].copy_(k_value, non_blocking=True)
            self.ev_in[g].record(self.copy_stream)
        compute.wait_event(self.ev_in[g])
        with torch.cuda.stream(compute):
            rec, _ = self.model.forward_record(self.img[g], self.img[g], self.kv[g], self.K[g])
            self.ev_free[g] = torch.cuda.Event()
            self.ev_free[g].record(compute)
            if self.post is not None:
                rec = self.post(rec)             # e.g. the multi-GPU gather of the packed records
            if self.host[s] is None or self.host[s].numel() != rec.numel():
                self.host[s] = torch.empty(rec.numel(), dtype=torch.float32).pin_memory()
            self.host[s].copy_(rec.reshape(-1), non_blocking=True)
            self.ev_done[s] = torch.cuda.Event()
            self.ev_done[s].record(compute)
        return s

    def result(self, ticket):
        self.ev_done[ticket].synchronize()
        if self.post is not None:
            return self.host[ticket]
        return dict(zip(capi.FIELD_NAMES, self.model._fields(self.host[ticket], self.offs, self.B)))


def get_rootNetwithRegInt_model(init_param_dict, args, device=None, precision="fp32"):
    """Factory with the reference's name and arguments (full_net.py:470-505); weights are loaded by the caller."""
    cfg = dict(args) if isinstance(args, dict) else {k: getattr(args, k) for k in dir(args) if not k.startswith("_")}
    keys = ("backbone_name", "rootnet_backbone_name", "n_iter", "rotation_dim", "reg_joint_map", "direct_reg_rot",
            "rot_iterative_matmul", "add_fc", "multi_kp", "kps_need_depth", "joint_conv_dim", "use_rpmg", "fix_root", "bbox_3d_shape",
            "reference_keypoint_id", "other_image_size")
    return HoliRobPoseB200(init_param_dict["robot_type"], {k: cfg[k] for k in keys if k in cfg}, device, precision)
