"""Shared helpers for the parity tests (inputs are re-derived from seeds recorded in the golden fixtures)."""
import os

import numpy as np
import torch

from hrp_b200 import consts, synth
from oracle import model as omodel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULLNET_CASES = [("panda", "resnet50"), ("kuka", "resnet50"), ("baxter", "resnet50"), ("panda", "hrnet32"),
                 ("baxter", "hrnet32")]
# north_star parity gates (fp32 / TF32 parity mode)
TOL_PX, TOL_RAD, TOL_DEPTH_M = 0.5, 1e-3, 1e-3

_oracles = {}


def oracle_for(robot, backbone, seed=1234):
    key = (robot, backbone, seed)
    if key not in _oracles:
        sd = synth.make_state_dict(robot, backbone, seed)
        _oracles[key] = (omodel.OracleModel(robot, sd, open(consts.urdf_path(robot)).read(), backbone), sd)
    return _oracles[key]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def inputs(B, seed):
    img, K, kv = synth.make_inputs(B, seed)
    return torch.from_numpy(img), torch.from_numpy(K), torch.from_numpy(kv)


def maxdiff(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
