#!/usr/bin/env python
"""BASELINE configs[4]: standalone batched FK + pinhole projection sweep, 1e4..1e7 poses, Panda / Kuka / Baxter.

Prints one JSON line per (robot, N): poses/s, achieved GB/s on the algorithmic bytes of SURVEY.md §8d
((dof+6+3+9)*4 in + nkpt*5*4 out per pose) and the fraction of the measured HBM copy bandwidth. `--check` also runs the
table interpreter (HRP_FK_GENERIC=1) on the same poses and reports the largest difference to the generated chain.
Under torchrun each rank sweeps its own shard (poses are independent) and rank 0 reports the aggregate: once without any
collective and, with `--gather`, once with the NCCL all-gather of the outputs (xyz + uv, nkpt*5*4 bytes per pose) inside
the timed region (SURVEY.md 8e asks for both; the roofline fraction refers to the run WITHOUT the gather).
`--cpu` adds the CPU row: the reference's own URDFRobot.get_keypoints[_root] (vectorised torch, lib/utils/urdf_robot.py:
95-135,193-223) timed up to 1e6 poses and its point_projection_from_3d_tensor Python loop (lib/utils/transforms.py:17-21)
timed up to 1e5 poses on this host's cores -- through oracle/refrun/harness.py from baseline/_ref when the copy is there,
else the oracle port; larger sizes are not run on the CPU (say so: nothing is extrapolated into the table).
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
    ap.add_argument("--gather", action="store_true", help="also time FK + NCCL all-gather of the outputs (N > 1)")
    ap.add_argument("--cpu", action="store_true", help="also time the reference algorithm on the host cores (rank 0)")
    args = ap.parse_args()
    # one JSON object per stdout line: whatever libraries print while initialising (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
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
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
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
            if args.gather and world > 1:
                nk = spec["nkpt"]
                gx = torch.empty(world * n_local, nk, 3, device=dev)
                gu = torch.empty(world * n_local, nk, 2, device=dev)

                def step():
                    x, u = fk.keypoints(q, rot, tr, K)
                    dist.all_gather_into_tensor(gx, x)
                    dist.all_gather_into_tensor(gu, u)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                dist.barrier()
                e0.record()
                for _ in range(args.iters):
                    step()
                e1.record()
                torch.cuda.synchronize()
                msg = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
                dist.all_reduce(msg, op=dist.ReduceOp.MAX)
                if rank == 0:
                    msg = float(msg)
                    print(json.dumps({"robot": robot, "poses": n_local * world, "n_gpus": world, "with_gather": True, "ms": msg,
                                      "poses_per_sec": n_local * world / msg * 1e3, "gathered_bytes_per_rank": n_local * world * nk * 20,
                                      "fk_only_ms": ms, "note": "FK + ncclAllGather of xyz and uv (every rank receives every pose's outputs)"}))
                del gx, gu
            del q, rot, tr, K
        if args.cpu and rank == 0:
            cpu_rows(robot, torch, synth, consts)
    if world > 1:
        dist.destroy_process_group()


def cpu_rows(robot, torch, synth, consts):
    """The reference algorithm on the host cores: FK (vectorised torch) to 1e6 poses, the projection loop to 1e5."""
    import time
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cwd = os.getcwd()
    kind = "port"
    fk_fn = proj_fn = None
    try:
        from oracle.refrun import harness
        if harness.available():
            ns = harness.setup()
            r = ns.urdf_robot.URDFRobot(robot)
            root = consts.ROBOTS[robot]["ref_kp"]
            fk_fn = (lambda q, rot, tr: r.get_keypoints(q, rot, tr)) if root == 0 else (lambda q, rot, tr: r.get_keypoints_root(q, rot, tr, root=root))
            proj_fn = ns.transforms.point_projection_from_3d_tensor
            kind = "reference"
    except Exception as e:
        print("fk cpu row: falling back to the oracle port (%s)" % e, file=sys.stderr)
    if fk_fn is None:
        from oracle import integral, kinematics, model as omodel
        om = omodel.OracleModel(robot, {}, open(consts.urdf_path(robot)).read())
        fk_fn = lambda q, rot, tr: torch.from_numpy(om.fk(q.numpy(), rot.numpy(), tr.numpy()))     # noqa: E731
        proj_fn = integral.project
    for n in (10_000, 100_000, 1_000_000):
        q, rot, tr, K = (torch.from_numpy(a) for a in synth.make_fk_inputs(robot, n, 3))
        with torch.no_grad():
            fk_fn(q[:1000], rot[:1000], tr[:1000])
            t0 = time.perf_counter()
            xyz = fk_fn(q, rot, tr)
            t_fk = time.perf_counter() - t0
            t_pr = None
            if n <= 100_000:
                t0 = time.perf_counter()
                proj_fn(K, xyz)
                t_pr = time.perf_counter() - t0
        print(json.dumps({"robot": robot, "poses": n, "cpu": True, "kind": kind, "cores": cores, "fk_ms": 1e3 * t_fk, "fk_poses_per_sec": n / t_fk,
                          "projection_ms": None if t_pr is None else 1e3 * t_pr, "projection_poses_per_sec": None if t_pr is None else n / t_pr,
                          "fk_plus_projection_poses_per_sec": None if t_pr is None else n / (t_fk + t_pr),
                          "note": "reference URDFRobot.get_keypoints[_root] (vectorised torch) + point_projection_from_3d_tensor (Python loop over poses); the loop is not run beyond 1e5 poses"}))
    os.chdir(cwd)


if __name__ == "__main__":
    main()
