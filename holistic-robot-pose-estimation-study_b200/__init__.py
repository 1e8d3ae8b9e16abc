"""HoliRobPose full-network inference forward for NVIDIA B200 (sm_100a).

Drop-in for the reference's `RootNetwithRegInt.forward(x_reg, x_root, k_value, K)` (lib/models/full_net.py:262-466)
plus the caller-side pinhole projection (lib/utils/transforms.py:17-21). Host side is Python/PyTorch (device memory,
streams, torch.distributed); all arithmetic runs in the hand-written CUDA library behind the C-ABI of
include/hrp_b200.h. There is no CPU fallback: importing `capi`/`model` without the built library raises.
"""
from . import arch, consts, synth, urdf  # noqa: F401  (pure-Python, no native code needed)

__all__ = ["arch", "consts", "synth", "urdf"]


def __getattr__(name):
    # native-backed modules are imported lazily so that pure-Python users (weight generation, URDF compilation)
    # do not need the CUDA library
    if name in ("capi", "model", "metrics"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name == "HoliRobPoseB200":
        from .model import HoliRobPoseB200
        return HoliRobPoseB200
    raise AttributeError(name)
