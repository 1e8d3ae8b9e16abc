#!/usr/bin/env python
"""Read an .ncu-rep of conv_tc_kernel: headline counters + stall samples aggregated around barrier/MMA/TMA instructions."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
want = ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed')
for i, h in enumerate(hdr):
    if h in want:
        print("%-70s %s %s" % (h, rows[1][i], [d[i] for d in data]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) > 10]
ci = {n: i for i, n in enumerate(h)}
base = int(data[0][0], 16)
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print("total samples", sum(int(r[ci["# Samples"]]) for r in data), "inst", sum(int(r[ci["Instructions Executed"]]) for r in data))
top = sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = sorted([(int(r[ci[n]]), n) for n in stall_cols], reverse=True)[:2]
    print("%05x %5s smp %8s inst  %-62s %s" % (int(r[0], 16) - base, r[ci["# Samples"]], r[ci["Instructions Executed"]], r[1].strip()[:62], st))
print("--- sync / async instructions")
for r in data:
    t = r[1]
    if any(k in t for k in ("TRYWAIT", "UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "BAR.SYNC", "ARRIVE", "UTCBAR", "NANOSLEEP")):
        print("%05x %5s smp %9s inst  %s" % (int(r[0], 16) - base, r[ci["# Samples"]], r[ci["Instructions Executed"]], t.strip()[:80]))
