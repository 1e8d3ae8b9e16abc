#!/usr/bin/env python
"""Summarise an HRP_DUMP_OPS csv (per-op device times of one un-graphed forward)."""
import collections, csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tot = sum(float(r['ms']) for r in rows)
print("total ms %.3f over %d ops" % (tot, len(rows)))
agg = collections.OrderedDict()
KIND = {0: "stem", 1: "conv", 2: "maxpool", 3: "fuse", 4: "avgpool", 5: "depth", 6: "rank", 7: "dec", 8: "softargmax", 9: "fk", 10: "stem_pack", 11: "block", 12: "chain"}
for r in rows:
    if r['kind'] in ('11', '12'):
        key = (KIND[int(r['kind'])], "%sx%s" % (r['Hi'], r['Hi']), r['Cin'] + "ch")
    elif r['kind'] != '1':
        key = (KIND[int(r['kind'])],)
    else:
        key = ("conv", "%sx%s" % (r['Hi'], r['Hi']), r['Cin'] + "->" + r['Cout'], "k" + r['k'], "s" + r['stride'], "res" + r['res'], "nchw" + r['nchw'], "cls" + r['cls'])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += float(r['ms']); a[2] = float(r['tflops'])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-70s n=%3d total %7.3f ms  avg %7.4f ms  %6.1f TF/s" % (" ".join(k), v[0], v[1], v[1] / v[0], v[2]))
if rows and "lane" in rows[0]:
    lanes = collections.OrderedDict()
    for r in rows:
        lanes[r["lane"]] = lanes.get(r["lane"], 0.0) + float(r["ms"])
    print("serial ms per lane:", {k: round(v, 3) for k, v in sorted(lanes.items(), key=lambda kv: int(kv[0]))})
