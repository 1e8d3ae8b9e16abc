#!/usr/bin/env python
"""SM time per forward from an ncu launch list taken at the multi-lane graph's caps (scripts/final_measure.sh):
duration x SMs held, per kernel family and for the heaviest launches. When four forwards are in flight and every launch is
capped at a quarter of the GPU the SMs are fully subscribed, so this sum / 148 is what a forward costs.
usage: sm_time.py launches.csv n_launches_of_one_forward"""
import collections, csv, re, sys
path, n = sys.argv[1], int(sys.argv[2])
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.reader(lines); hdr = next(rd); ci = {k: i for i, k in enumerate(hdr)}
per = collections.OrderedDict()
for r in rd:
    d = per.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]]})
    v = float(r[ci["Metric Value"]].replace(",", "")); u = r[ci["Metric Unit"]]
    d[r[ci["Metric Name"]]] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
rows, tot = [], 0.0
for i in sorted(per)[-n:]:
    d = per[i]
    g = int(d["grid"].strip("()").split(",")[0]); blk = int(d["block"].strip("()").split(",")[0])
    nm = re.sub(r".*::", "", re.sub(r"\(.*", "", d["name"]).replace("void ", ""))
    # conv_tc runs two 320-thread CTAs per SM when its grid is 74 at the quarter-GPU cap; everything else one CTA per SM (or a
    # grid of small CTAs that spreads over all 148)
    sms = 37 if ("conv_tc" in nm and g == 74) else min(g, 148)
    t = d["gpu__time_duration.sum"]
    rows.append((t * sms, nm, g, blk, t, d.get("lts__t_bytes.sum", 0) / 1e6, (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6,
                 d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)))
    tot += t * sms
print("one forward of 64 frames, %d launches: %.0f SM-us = %.2f ms of the whole GPU (148 SMs)" % (len(rows), tot, tot / 148e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    agg[r[1]][0] += 1; agg[r[1]][1] += r[0]
print("\n%-36s %5s %10s %7s" % ("kernel", "n", "SM-us", "share"))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %5d %10.0f %6.1f%%" % (k[:36], c, v, 100 * v / tot))
print("\nheaviest launches:\n%-30s %5s %6s %9s %8s %8s %8s %8s" % ("kernel", "grid", "block", "us", "SM-us", "L2 MB", "DRAM MB", "tensor%"))
for r in sorted(rows, reverse=True)[:30]:
    print("%-30s %5d %6d %9.1f %8.0f %8.0f %8.0f %8.1f" % (r[1][:30], r[2], r[3], r[4], r[0], r[5], r[6], r[7]))
