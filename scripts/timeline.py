#!/usr/bin/env python
"""Development aid: run a few graph replays with HRP_TIMELINE set and summarise per-lane busy time of the last replay."""
import os, sys, csv, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1]
os.environ["HRP_TIMELINE"] = out
import torch
import hrp_b200  # noqa
from hrp_b200 import synth
from hrp_b200.model import HoliRobPoseB200
dev = torch.device("cuda", 0)
m = HoliRobPoseB200("panda", {"backbone_name": "resnet50"}, device=dev, precision="bf16")
m.load_state_dict(synth.make_state_dict("panda", "resnet50"))
img, K, kv = (torch.from_numpy(a).to(dev) for a in synth.make_inputs(64, 1))
for _ in range(5):
    m.forward_dict(img, K, kv)
torch.cuda.synchronize()
del m
import gc; gc.collect()
rows = list(csv.DictReader(open(out)))
end = max(int(r["end_ns"]) for r in rows)
print("replay span %.3f ms over %d ops" % (end / 1e6, len(rows)))
lanes = collections.OrderedDict()
for r in rows:
    l = int(r["lane"]); s, e = int(r["start_ns"]), int(r["end_ns"])
    d = lanes.setdefault(l, [0, 1 << 62, 0, 0])
    d[0] += e - s; d[1] = min(d[1], s); d[2] = max(d[2], e); d[3] += 1
for l, (busy, s, e, n) in sorted(lanes.items()):
    print("lane %d: %3d ops, busy %.3f ms, first start %.3f ms, last end %.3f ms" % (l, n, busy / 1e6, s / 1e6, e / 1e6))
# concurrency profile: how many ops are in flight over time (sampled)
ev = []
for r in rows:
    ev.append((int(r["start_ns"]), 1)); ev.append((int(r["end_ns"]), -1))
ev.sort()
cur, last, hist = 0, 0, collections.Counter()
for t, d in ev:
    hist[cur] += t - last; last = t; cur += d
print("time with k ops in flight (ms):", {k: round(v / 1e6, 3) for k, v in sorted(hist.items())})
# per (kind, lane) mean duration
agg = collections.OrderedDict()
for r in rows:
    k = (int(r["kind"]), int(r["lane"]), int(r["Hi"]), int(r["Cin"]), int(r["Cout"]), int(r["k"]), int(r["stride"]))
    a = agg.setdefault(k, [0, 0])
    a[0] += 1; a[1] += int(r["end_ns"]) - int(r["start_ns"])
print("kind lane H Cin Cout k s : n mean_us total_ms")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(k, n, round(t / n / 1e3, 1), round(t / 1e6, 3))
