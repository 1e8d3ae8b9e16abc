"""ctypes binding of libhrp_b200.so (C ABI: include/hrp_b200.h). No CPU fallback: a missing library raises."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhrp_b200.so")

NUM_FIELDS = 11          # the ten named fields + HRP_F_DEPTHS (multi_kp: every regressed depth; width 0 otherwise)
NUM_CLASSES = 7
FIELD_NAMES = ("joint_angles", "rot6d", "trans", "root_uv", "root_depth", "uvd", "kp3d_int", "kp3d_fk", "kp2d_int",
               "kp2d_fk")
CLASS_NAMES = ("conv_tensor", "conv_fp32", "stem", "elementwise", "heads", "softargmax", "fk")
PREC = {"fp32": 0, "tf32": 1, "bf16": 2, "tf32x3": 3, "f16": 4}
BACKBONE = {"resnet": 0, "resnet50": 0, "hrnet": 1, "hrnet32": 1}

EXPORTS = (
    "hrp_fk_create", "hrp_fk_destroy", "hrp_fk_project", "hrp_softargmax3d_workspace", "hrp_softargmax3d",
    "hrp_conv2d_nhwc", "hrp_basic_block_nhwc", "hrp_basic_chain_nhwc", "hrp_create", "hrp_destroy", "hrp_num_weights", "hrp_weight_name", "hrp_weight_shape",
    "hrp_set_weight", "hrp_finalize_weights", "hrp_output_offsets", "hrp_workspace_bytes", "hrp_forward", "hrp_forward_ex",
    "hrp_forward_timed", "hrp_release_plans", "hrp_forward_u8", "hrp_crop_resize_u8",
    "hrp_metrics_batch", "hrp_summary_workspace", "hrp_summary_add_pck",
    "hrp_p2p_create", "hrp_p2p_handle", "hrp_p2p_connect", "hrp_p2p_all_gather", "hrp_p2p_status", "hrp_p2p_destroy",
    "hrp_set_option", "hrp_launch_count", "hrp_debug_tensor", "hrp_forward_profile", "hrp_conv_bench", "hrp_last_error",
    "hrp_version",
)


class FkProgram(C.Structure):
    _fields_ = [("dof", C.c_int32), ("nkpt", C.c_int32), ("n_steps", C.c_int32), ("n_slots", C.c_int32),
                ("root_kp", C.c_int32), ("root_step", C.c_int32),
                ("step_type", C.POINTER(C.c_int32)), ("step_parent", C.POINTER(C.c_int32)),
                ("step_save", C.POINTER(C.c_int32)), ("step_q", C.POINTER(C.c_int32)),
                ("step_mul", C.POINTER(C.c_float)), ("step_off", C.POINTER(C.c_float)),
                ("step_origin", C.POINTER(C.c_float)), ("step_axis", C.POINTER(C.c_float)),
                ("kp_step", C.POINTER(C.c_int32)), ("kp_index", C.POINTER(C.c_int32)),
                ("kp_offset", C.POINTER(C.c_float)), ("root_fixed", C.POINTER(C.c_float))]


class Config(C.Structure):
    _fields_ = [("backbone", C.c_int32), ("precision", C.c_int32), ("n_iter", C.c_int32), ("fix_root", C.c_int32),
                ("image_size", C.c_float), ("depth_factor", C.c_float),
                ("direct_reg_rot", C.c_int32), ("rot_iterative_matmul", C.c_int32), ("add_fc", C.c_int32),
                ("depth_num", C.c_int32), ("depth_root", C.c_int32), ("reg_joint_map", C.c_int32),
                ("joint_conv_dim", C.c_int32 * 3), ("joint_bounds", C.c_float * 32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libhrp_b200.so is not built (%s); run ./build.sh or __graft_entry__.build(). "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
        L.hrp_last_error.restype = C.c_char_p
        L.hrp_version.restype = C.c_char_p
        L.hrp_fk_create.argtypes = [C.POINTER(FkProgram), C.POINTER(vp)]
        L.hrp_fk_destroy.argtypes = [vp]
        L.hrp_fk_destroy.restype = None
        L.hrp_fk_project.argtypes = [vp, f32p, f32p, f32p, f32p, i64, f32p, f32p, vp]
        L.hrp_softargmax3d_workspace.argtypes = [i32] * 5
        L.hrp_softargmax3d_workspace.restype = C.c_size_t
        L.hrp_softargmax3d.argtypes = [f32p, i32, i32, i32, i32, i32, f32p, f32p, C.c_float, C.c_float, i32, i32, f32p,
                                       f32p, vp, C.c_size_t, vp]
        L.hrp_conv2d_nhwc.argtypes = [f32p] * 5 + [i32] * 11 + [vp]
        L.hrp_create.argtypes = [C.POINTER(Config), C.POINTER(FkProgram), i32, C.POINTER(vp)]
        L.hrp_destroy.argtypes = [vp]
        L.hrp_destroy.restype = None
        L.hrp_num_weights.argtypes = [vp]
        L.hrp_weight_name.argtypes = [vp, i32]
        L.hrp_weight_name.restype = C.c_char_p
        L.hrp_weight_shape.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i32)]
        L.hrp_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32, i32]
        L.hrp_finalize_weights.argtypes = [vp]
        L.hrp_output_offsets.argtypes = [vp, i32, C.POINTER(i64)]
        L.hrp_workspace_bytes.argtypes = [vp, i32]
        L.hrp_workspace_bytes.restype = C.c_size_t
        L.hrp_forward.argtypes = [vp, f32p, f32p, f32p, f32p, i32, f32p, vp]
        L.hrp_forward_ex.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, f32p, i32, f32p, vp]
        L.hrp_forward_timed.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, f32p, i32, f32p, C.POINTER(C.c_float), vp]
        L.hrp_release_plans.argtypes = [vp]
        L.hrp_forward_u8.argtypes = [vp, f32p, f32p, f32p, f32p, i32, f32p, vp]
        L.hrp_crop_resize_u8.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.hrp_metrics_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        L.hrp_summary_workspace.argtypes = [i64]
        L.hrp_summary_workspace.restype = C.c_size_t
        L.hrp_summary_add_pck.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
        L.hrp_p2p_create.argtypes = [i32, i32, C.c_size_t, i32, C.POINTER(vp)]
        L.hrp_p2p_handle.argtypes = [vp, vp]
        L.hrp_p2p_connect.argtypes = [vp, vp]
        L.hrp_p2p_all_gather.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.hrp_p2p_status.argtypes = [vp]
        L.hrp_p2p_destroy.argtypes = [vp]
        L.hrp_p2p_destroy.restype = None
        L.hrp_set_option.argtypes = [vp, C.c_char_p, i64]
        L.hrp_launch_count.argtypes = [vp]
        L.hrp_launch_count.restype = i64
        L.hrp_debug_tensor.argtypes = [vp, C.c_char_p, i32, f32p, C.POINTER(i64), vp]
        L.hrp_forward_profile.argtypes = [vp, f32p, f32p, f32p, f32p, i32, f32p, C.POINTER(C.c_float), C.POINTER(i64),
                                          C.POINTER(C.c_double), vp]
        L.hrp_basic_block_nhwc.argtypes = [f32p] * 6 + [i32] * 4 + [vp]
        L.hrp_basic_chain_nhwc.argtypes = [f32p] * 3 + [i32, f32p] + [i32] * 4 + [vp]
        L.hrp_conv_bench.argtypes = [i32] * 10 + [C.POINTER(C.c_float), vp]
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError("hrp_b200: %s (status %d)" % (lib().hrp_last_error().decode(), status))


class _Keep:
    """ctypes struct + the numpy arrays its pointers refer to."""


def fk_program_struct(prog):
    """urdf.KinematicProgram -> (FkProgram, keep-alive holder)."""
    arrs = prog.arrays()
    keep = _Keep()
    keep.arrs = {k: np.ascontiguousarray(v) for k, v in arrs.items()}

    def ip(k):
        return keep.arrs[k].ctypes.data_as(C.POINTER(C.c_int32))

    def fp(k):
        return keep.arrs[k].ctypes.data_as(C.POINTER(C.c_float))

    s = FkProgram(prog.dof, prog.nkpt, len(prog.step_type), prog.n_slots, prog.root_kp, prog.root_step,
                  ip("step_type"), ip("step_parent"), ip("step_save"), ip("step_q"), fp("step_mul"), fp("step_off"),
                  fp("step_origin"), fp("step_axis"), ip("kp_step"), ip("kp_index"), fp("kp_offset"), fp("root_fixed"))
    keep.struct = s
    return s, keep
