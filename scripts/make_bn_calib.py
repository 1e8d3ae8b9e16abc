"""Produce data/bn_calib_v2.npz: BatchNorm running statistics for the synthetic weights (one-off, CPU, ~1 min).

For each weight seed and each sub-network (ResNet-50 + deconv keypoint branch, HRNet-W32 keypoint branch, HRNet-W32
DepthNet) run the oracle forward once over CALIB_BATCH seeded noise images with BN in batch-statistics mode and store
(mean, unbiased var) per BN layer. Keys: "<seed>/<resnet50|hrnet32|rootnet>/<state-dict name>".
usage: make_bn_calib.py [damped|undamped]  (undamped -> data/bn_calib_undamped.npz, see synth.make_state_dict)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hrp_b200  # noqa: E402
from hrp_b200 import consts, synth  # noqa: E402
from oracle import network  # noqa: E402

SEEDS = (1234,)


def main(recipe="damped"):
    torch.manual_seed(0)
    out = {}
    x = torch.from_numpy(synth.make_images(synth.CALIB_BATCH, synth.CALIB_IMAGE_SEED))
    for seed in SEEDS:
        for variant in ("resnet50", "hrnet32"):
            sd = {k: torch.from_numpy(np.asarray(v)) for k, v in
                  synth.make_state_dict("panda", variant, seed, calibrated=False, recipe=recipe).items()}
            with torch.no_grad():
                c = network.Calib()
                if variant == "resnet50":
                    f = network.resnet50(x, sd, "reg_backbone.", c)
                    network.deconv_head(f, sd, c)
                    r = network.Calib()
                    network.hrnet_w32(x, sd, "rootnet_backbone.", False, r)
                    for k, v in r.stats.items():
                        out["%d/rootnet/%s" % (seed, k)] = v.numpy()
                else:
                    network.hrnet_w32(x, sd, "reg_backbone.", True, c)
                for k, v in c.stats.items():
                    out["%d/%s/%s" % (seed, variant, k)] = v.numpy()
            print(seed, variant, len(out))
    path = synth.CALIB_FILE if recipe == "damped" else synth.CALIB_FILE_UNDAMPED
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "damped")
