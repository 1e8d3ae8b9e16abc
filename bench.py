#!/usr/bin/env python
"""Headline benchmark: Panda full-network inference forward, frames/sec @256x256 (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|tf32|bf16] [--backbone resnet50|hrnet32]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, NCCL)
  python bench.py --impl reference ...      (the reference algorithm on the host CPU cores: oracle port, see below)

A step is one forward of the whole path (DepthNet HRNet-W32 + keypoint backbone + heatmap soft-argmax + heads + FK +
both projections) over one batch of 64 synthetic frames per GPU with calibrated random-init weights. `value` is timed
with the inputs already resident in HBM (rotating over distinct input batches whose total size exceeds L2); `e2e` is the
same metric through the public Python API with pinned HOST buffers, host->device and device->host copies inside the
timed region. Batch shards are independent; with N>1 the only collective is the all-gather of the packed output
records (inside the timed region). Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.

Reference arm: the reference is pure Python on PyTorch and cannot travel to the GPU box (nor run without its
unavailable dependencies), so `--impl reference` times the oracle port (oracle/model.py: the same torch CPU ops the
reference dispatches, pinned bit-for-bit to the reference's outputs in this repo's golden fixtures) on all host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROBOT = "panda"
BATCH_PER_GPU = 64
WEIGHT_SEED = 1234


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_fps(torch, frames_per_step, steps, warmup, backbone):
    """The oracle port (reference algorithm, PyTorch CPU fp32) on all host cores; returns (fps, cores, ms/step)."""
    import hrp_b200  # noqa: F401
    from hrp_b200 import consts, synth
    from oracle import model as omodel
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(ROBOT, backbone, WEIGHT_SEED)
    om = omodel.OracleModel(ROBOT, sd, open(consts.urdf_path(ROBOT)).read(), backbone)
    img, K, kv = (torch.from_numpy(a) for a in synth.make_inputs(frames_per_step, 31337))
    for _ in range(warmup):
        om.forward_dict(img, img, kv, K)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        om.forward_dict(img, img, kv, K)
        times.append(time.perf_counter() - t0)
    tot = sum(times)
    return frames_per_step * steps / tot, cores, 1e3 * tot / steps


def main():
    global ROBOT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HRP_PRECISION", "bf16"), choices=["fp32", "tf32", "bf16"],
                    help="conv/linear contraction arithmetic: bf16 = throughput mode (default), fp32 = parity mode")
    ap.add_argument("--backbone", default="resnet50", choices=["resnet50", "hrnet32"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="frames per GPU per step")
    ap.add_argument("--robot", default=ROBOT, choices=["panda", "kuka", "baxter"],
                    help="panda = the headline workload (BASELINE configs[1]); kuka / baxter with --batch 256 / 128 are configs[2] / [3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-families", action="store_true", help="skip the tf32 / fp32 family lines")
    args = ap.parse_args()
    ROBOT = args.robot

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    workload = ROBOT.capitalize() + " full network (%s keypoint backbone + HRNet-W32 DepthNet + heatmap soft-argmax + heads + FK/projection), batch %d per GPU, 256x256 synthetic RGB" % (
        "ResNet-50+deconv" if args.backbone == "resnet50" else "HRNet-W32", args.batch)

    # the contract is ONE JSON line on stdout: anything a library prints while we set up (NCCL's version banner goes to
    # stdout) is sent to stderr instead, and the real stdout comes back just before the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    import torch

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = 4
        fps, cores, ms = cpu_reference_fps(torch, sample, args.steps, args.warmup, args.backbone)
        line = {"impl": "reference", "metric": "%s_fullnet_frames_per_sec" % ROBOT, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "l2": "n/a (CPU)"},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d frames per step of the batch-%d workload, oracle port of the reference (torch %s CPU fp32)" % (sample, args.batch, torch.__version__)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch.distributed as dist
    import hrp_b200  # noqa: F401
    from hrp_b200 import arch, capi, consts, synth, dist as hdist
    from hrp_b200.model import HoliRobPoseB200, FkRobot, HostPipeline, soft_argmax

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B = args.batch
    spec = consts.ROBOTS[ROBOT]

    model = HoliRobPoseB200(ROBOT, {"backbone_name": args.backbone}, device=dev, precision=args.precision)
    model.load_state_dict(synth.make_state_dict(ROBOT, args.backbone, WEIGHT_SEED))

    # distinct input batches, rotated so consecutive steps never re-read the same images from L2 (4 x 50 MB > 126 MB L2)
    NSETS = 4
    sets_host, sets_dev = [], []
    for s in range(NSETS):
        img, K, kv = synth.make_inputs(B, 5000 + 100 * rank + s)
        h = [torch.from_numpy(a).pin_memory() for a in (img, K, kv)]
        sets_host.append(h)
        sets_dev.append([t.to(dev) for t in h])
    offs = model._record(B, dev)
    rec_bytes = offs[-1] * 4

    # consecutive forwards go to two streams: the library alternates between two plans (workspace + graph) per batch
    # size, so the low-parallelism tail of step i overlaps the head of step i+1 (a stream of independent batches)
    side = [torch.cuda.Stream(dev) for _ in range(int(os.environ.get("HRP_SLOTS", "3")))]

    def step_resident(i):
        img, K, kv = sets_dev[i % NSETS]
        with torch.cuda.stream(side[i % len(side)]):
            rec, _ = model.forward_record(img, img, kv, K)
            if world > 1:
                hdist.gather_records(rec, B, spec["dof"], spec["nkpt"])
        return rec

    # e2e: the public streaming API (HostPipeline): every step uploads its own batch from pinned host memory and reads its
    # own result record back; with two slots the upload of step i+1 overlaps the forward of step i
    def gathered(rec):
        if world > 1:
            hdist.gather_records(rec, B, spec["dof"], spec["nkpt"])
        return rec
    pipe = HostPipeline(model, B, depth=int(os.environ.get("HRP_SLOTS", "3")), post=gathered)
    pending = []

    def step_e2e(i):
        img, K, kv = sets_host[i % NSETS]
        pending.append(pipe.submit(img, K, kv))
        if len(pending) >= pipe.depth:
            pipe.result(pending.pop(0))                  # the caller consumes every step's record

    def drain_e2e():
        while pending:
            pipe.result(pending.pop(0))

    def timed(fn, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for st_ in side:
            st_.wait_event(e0)
        for i in range(steps):
            fn(i)
        if fn is step_e2e:
            drain_e2e()
        for st_ in side:
            torch.cuda.current_stream().wait_stream(st_)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for i in range(warmup):
        step_resident(i)
    for i in range(3):
        step_e2e(i)
    drain_e2e()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(step_resident, args.steps)
    sampler.stop_flag = True
    sampler.join(1.0)
    launches = model.launch_count() * args.steps
    ms_e2e = timed(step_e2e, args.steps)

    frames = B * world * args.steps
    fps = frames / (ms_total * 1e-3)
    fps_e2e = frames / (ms_e2e * 1e-3)
    flops_frame = arch.flops_per_frame(ROBOT, args.backbone)

    # ---- roofline of the dominant kernel family (convolutions), measured live: CUDA events around every launch ----------
    img, K, kv = sets_dev[0]
    prof = model.profile(img, img, kv, K)
    conv = prof["conv_tensor"] if prof["conv_tensor"]["launches"] else prof["conv_fp32"]
    conv_name = "conv_tensor" if prof["conv_tensor"]["launches"] else "conv_fp32"
    total_ms = sum(v["ms"] for v in prof.values())
    serial_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    tensor_peak = peaks["bf16_sustained"] * (0.5 if args.precision == "tf32" else 1.0)
    # The conv family is ~98 % of the step's FLOPs and its launches overlap across the graph's lanes, so per-launch
    # durations are not additive: achieved = the family's algorithmic FLOPs of one step / the measured step time (CUDA
    # events around the timed graph replays). `serial_launch_tflops` is the same FLOPs / the SUM of per-launch durations of
    # one un-graphed, single-stream forward (CUDA events around every launch).
    step_s = ms_total / args.steps * 1e-3
    conv_tflops = conv["flops"] / step_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_dram_traffic_per_step.json")
    if os.path.exists(tpath) and args.precision == "bf16" and B == BATCH_PER_GPU:
        traffic = json.load(open(tpath)).get("dram_bytes_per_step")
    roofline = {"bound": "tensor", "kernel": "%s (conv_tc / conv_slab / conv_block / conv_chain kernels, %d launches per step)" % (conv_name, conv["launches"]),
                "achieved": conv_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": conv_tflops / tensor_peak, "traffic": traffic,
                "traffic_note": "DRAM bytes of ALL kernels of one step (ncu dram__bytes_read+write, profiles/r01_launches_bf16_b64.csv); null when not profiled for this config",
                "peak_source": "%s bf16 sustained%s" % (peaks["source"], " / 2 (tf32)" if args.precision == "tf32" else ""),
                "flops_per_launch_avg": conv["flops"] / max(conv["launches"], 1),
                "launches_per_step": conv["launches"], "serial_launch_tflops": serial_tflops,
                "share_of_serial_step": conv["ms"] / total_ms if total_ms else None,
                "by_class_serial_ms": {k: round(v["ms"], 4) for k, v in prof.items()}}

    # ---- the two HBM-bound kernels in isolation, on inputs larger than L2 ------------------------------------------------
    extra = {}
    if rank == 0:
        def ev_time(fn, n):
            fn(); fn(); fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        nk = spec["nkpt"]
        hm = torch.randn(B, nk * 64, 64, 64, device=dev)
        Kd, kvd = sets_dev[0][1], sets_dev[0][2]
        rz = torch.ones(B, device=dev)
        ms = ev_time(lambda: soft_argmax(hm, nk, Kd, rz, 1.3, 256.0, spec["ref_kp"], True), 20)
        nbytes = hm.numel() * 4 + B * nk * 24
        extra["softargmax"] = {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"], "bytes_per_launch": nbytes, "ms": ms,
                               "shape": "B=%d K=%d 64^3 fp32" % (B, nk)}
        del hm
        n = 10_000_000
        fk = FkRobot(ROBOT)
        q, rot, tr, Kf = (torch.from_numpy(a).to(dev) for a in synth.make_fk_inputs(ROBOT, 100_000, 1))
        rep = n // 100_000
        q, rot, tr, Kf = q.repeat(rep, 1), rot.repeat(rep, 1), tr.repeat(rep, 1), Kf.repeat(rep, 1, 1)
        ms = ev_time(lambda: fk.keypoints(q, rot, tr, Kf), 10)
        per_pose = (spec["dof"] + 6 + 3 + 9) * 4 + nk * 5 * 4
        extra["fk"] = {"bound": "hbm", "achieved": n * per_pose / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": n * per_pose / ms / 1e6 / peaks["hbm_gbs"], "bytes_per_pose": per_pose, "poses": n, "ms": ms,
                       "poses_per_sec": n / ms * 1e3}
        del q, rot, tr, Kf

    # ---- the other precision families on the same workload (single GPU only; short, same timing rules) ------------------
    families = {}
    ws_gb = capi.lib().hrp_workspace_bytes(model._h, B) / 2 ** 30
    if rank == 0 and world == 1 and not args.no_families:
        del model
        torch.cuda.empty_cache()
        for prec in ("tf32", "fp32"):
            if prec == args.precision:
                continue
            m2 = HoliRobPoseB200(ROBOT, {"backbone_name": args.backbone}, device=dev, precision=prec)
            m2.load_state_dict(synth.make_state_dict(ROBOT, args.backbone, WEIGHT_SEED))
            def fam_step(i):
                st_ = side[i % len(side)]
                with torch.cuda.stream(st_):
                    m2.forward_record(sets_dev[i % NSETS][0], sets_dev[i % NSETS][0], sets_dev[i % NSETS][2], sets_dev[i % NSETS][1])
            for i in range(2 * len(side)):
                fam_step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n2 = 6 if prec == "fp32" else 12
            e0.record()
            for st_ in side:
                st_.wait_event(e0)
            for i in range(n2):
                fam_step(i)
            for st_ in side:
                torch.cuda.current_stream().wait_stream(st_)
            e1.record()
            torch.cuda.synchronize()
            families[prec] = {"value": B * n2 / (e0.elapsed_time(e1) * 1e-3), "unit": "frames/s", "steps": n2,
                              "parity": {"tf32": "north_star gates (1e-3 rad, 1 mm, 0.5 px) on this configuration",
                                         "fp32": "north_star gates on every configuration"}[prec]}
            del m2
            torch.cuda.empty_cache()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cfps, cores, cms = cpu_reference_fps(torch, 1, 10, 2, args.backbone)
        cpu_base = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                    "sample": "batch 1 x 10 forwards (+2 warm-up) of the oracle port of the reference, torch %s CPU fp32, %.0f ms/frame" % (torch.__version__, cms)}

    if rank == 0:
        line = {"metric": "%s_fullnet_frames_per_sec" % ROBOT, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": {"workload": workload, "robot": ROBOT, "backbone": args.backbone, "precision": args.precision,
                           "batch_per_gpu": B, "global_batch": B * world, "gflop_per_frame": flops_frame / 1e9,
                           "l2": "inputs rotate over %d distinct batches (%d MB > L2); the %.1f GB activation workspace is rewritten every step" % (
                               NSETS, NSETS * B * 3 * 256 * 256 * 4 // 2 ** 20, ws_gb),
                           "parallelism": "batch-sharded x%d, NCCL all-gather of output records" % world if world > 1 else "single GPU",
                           "pipelining": "consecutive steps are enqueued round-robin on %d streams (the library keeps as many plans per batch size), so the tail of step i overlaps the head of step i+1; every step is a complete forward of its own batch" % len(side),
                           "weights": "calibrated random init, seed %d" % WEIGHT_SEED},
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": B * (3 * 256 * 256 + 9 + 1) * 4,
                        "d2h_bytes_per_step": rec_bytes, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "rooflines_hbm": extra,
                "families": families, "cpu_baseline": cpu_base}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
