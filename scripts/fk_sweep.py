#!/usr/bin/env python
"""BASELINE configs[4]: standalone batched FK + pinhole projection sweep, 1e4..1e7 poses, Panda / Kuka / Baxter.

Prints one JSON line per (robot, N): poses/s, achieved GB/s on the algorithmic bytes of SURVEY.md §8d
((dof+6+3+9)*4 in + nkpt*5*4 out per pose) and the fraction of the measured HBM copy bandwidth. `--check` also runs the
table interpreter (HRP_FK_GENERIC=1) on the same poses and reports the largest difference to the generated chain.
Under torchrun each rank sweeps its own shard (no collective: poses are independent) and rank 0 reports the aggregate.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", default="panda,kuka,baxter")
    ap.add_argument("--sizes", default="10000,100000,1000000,10000000")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import hrp_b200  # noqa: F401
    from hrp_b200 import consts, synth
    from hrp_b200.model import FkRobot
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    for robot in args.robots.split(","):
        spec = consts.ROBOTS[robot]
        fk = FkRobot(robot)
        per_pose = (spec["dof"] + 6 + 3 + 9) * 4 + spec["nkpt"] * 5 * 4
        base = [torch.from_numpy(a).to(dev) for a in synth.make_fk_inputs(robot, 100_000, 1 + rank)]
        if args.check:
            os.environ["HRP_FK_GENERIC"] = "1"
            fk_gen = FkRobot(robot)
            del os.environ["HRP_FK_GENERIC"]
            x0, u0 = fk.keypoints(*base)
            x1, u1 = fk_gen.keypoints(*base)
            print(json.dumps({"robot": robot, "check": "generated chain vs table interpreter, 1e5 poses",
                              "max_abs_xyz_m": float((x0 - x1).abs().max()), "max_abs_uv_px": float((u0 - u1).abs().max())}))
        for n in [int(s) for s in args.sizes.split(",")]:
            n_local = n // world
            rep = max(1, (n_local + 99_999) // 100_000)
            q, rot, tr, K = (t.repeat(rep, *([1] * (t.dim() - 1)))[:n_local].contiguous() for t in base)
            for _ in range(3):
                fk.keypoints(q, rot, tr, K)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fk.keypoints(q, rot, tr, K)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ms = float(ms)
            if rank == 0:
                gbs = n_local * world * per_pose / ms / 1e6
                print(json.dumps({"robot": robot, "poses": n_local * world, "n_gpus": world, "ms": ms, "poses_per_sec": n_local * world / ms * 1e3,
                                  "bytes_per_pose": per_pose, "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / (peaks["hbm_gbs"] * world),
                                  "note": "includes two torch.empty output allocations per call; N <= 1e5 is L2/launch-bound"}))
            del q, rot, tr, K
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
