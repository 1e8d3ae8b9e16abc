#!/usr/bin/env python
"""Summarise an ncu launch list (--csv, metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum and
optionally sm__pipe_tensor_cycles_active / dram__throughput / sm__throughput percentages) over its LAST n launches (one
graph replay): python scripts/ncu_launches.py launches.csv n [out.json]"""
import collections, csv, json, re, sys
path, n = sys.argv[1], int(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ci = {k: i for i, k in enumerate(hdr)}
per = collections.OrderedDict()
for r in rd:
    d = per.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]]})
    v = float(r[ci["Metric Value"]].replace(",", ""))
    u = r[ci["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[ci["Metric Name"]]] = v
ids = sorted(per)[-n:]
def short(nm):
    nm = re.sub(r"\(.*$", "", nm)
    nm = re.sub(r"^void ", "", nm)
    return re.sub(r"(hrp::)?(<?unnamed>::|\(anonymous namespace\)::)", "", nm)
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for i in ids:
    d = per[i]
    a = agg[short(d["name"])]
    t = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["us"] += t
    a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["l2"] += d.get("lts__t_bytes.sum", 0.0)
    for k, s in (("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dramp"),
                 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm")):
        if k in d: a[s] += d[k] * t
tot = sum(a["us"] for a in agg.values())
print("last %d launches: %.1f us of launch time, %.3f GB of DRAM traffic, %.3f GB of L2 traffic" % (
    len(ids), tot, sum(a["dram"] for a in agg.values()) / 1e9, sum(a["l2"] for a in agg.values()) / 1e9))
print("%-46s %5s %10s %7s %10s %10s %8s" % ("kernel", "n", "total us", "share", "dram MB", "L2 MB", "tensor%"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print("%-46s %5d %10.1f %6.1f%% %10.1f %10.1f %8.1f" % (k[:46], a["n"], a["us"], 100 * a["us"] / tot, a["dram"] / 1e6, a["l2"] / 1e6, a["tensor"] / max(a["us"], 1e-9)))
if len(sys.argv) > 3:
    ks = sorted(agg, key=lambda k: -agg[k]["us"])
    json.dump({"dram_bytes_per_step": sum(a["dram"] for a in agg.values()), "l2_bytes_per_step": sum(a["l2"] for a in agg.values()),
               "launches": len(ids), "sum_of_launch_durations_us": round(tot, 3),
               "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,... --clock-control none over scripts/one_step.py (f16, batch 64), last forward (2 pre-graph launches + the graph's nodes); cold-cache (ncu flushes caches between launches), serialised launches, each launch at the graph's quarter-GPU cap: compare shares, not absolutes",
               "by_kernel_l2_mb": {k: round(agg[k]["l2"] / 1e6, 1) for k in sorted(agg, key=lambda k: -agg[k]["us"])},
               "by_kernel_us": {k: round(agg[k]["us"], 1) for k in ks}, "by_kernel_launches": {k: int(agg[k]["n"]) for k in ks},
               "by_kernel_dram_mb": {k: round(agg[k]["dram"] / 1e6, 1) for k in ks}}, open(sys.argv[3], "w"), indent=1)
