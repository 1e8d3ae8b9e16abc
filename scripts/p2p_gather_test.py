#!/usr/bin/env python
"""Peer-memory output gather (csrc/p2p_gather.cu) against NCCL's all-gather on the same records, and both timed:
torchrun --nproc-per-node N scripts/p2p_gather_test.py [floats per record ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import hrp_b200  # noqa
from hrp_b200 import dist as hd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_so, _fd = sys.stdout, os.dup(1)
os.dup2(2, 1)                                                 # NCCL's banner goes to stderr
dist.init_process_group("nccl", device_id=dev)
sizes = [int(v) for v in sys.argv[1:]] or [7104, 7101, 63488, 1]
res = []
for numel in sizes:
    pg = hd.PeerGather(numel, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(4)]
    bad = 0
    for step in range(64):                                       # rotating streams like the bench: gathers of different steps overlap
        with torch.cuda.stream(streams[step % 4]):
            rec = torch.arange(numel, device=dev, dtype=torch.float32) * 0.5 + 1000.0 * rank + step
            got = pg.all_gather(rec)
            want = torch.empty(world * numel, device=dev, dtype=torch.float32)
            dist.all_gather_into_tensor(want, rec)
            bad += int(not torch.equal(got.reshape(-1), want))
    torch.cuda.synchronize()
    pg.check()
    # timing on one stream, device time, max over ranks
    rec = torch.randn(numel, device=dev)
    def timed(fn, n=200):
        for _ in range(20):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) * 1e3
    out = torch.empty(world * numel, device=dev)
    us_p2p = timed(lambda: pg.all_gather(rec))
    us_nccl = timed(lambda: dist.all_gather_into_tensor(out, rec))
    pg.check()
    res.append({"floats_per_rank": numel, "world": world, "mismatching_steps": bad, "p2p_us": round(us_p2p, 2), "nccl_us": round(us_nccl, 2)})
    pg.close()
os.dup2(_fd, 1)
if rank == 0:
    for r in res:
        print(json.dumps(r))
dist.destroy_process_group()
