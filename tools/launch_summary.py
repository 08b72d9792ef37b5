"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total time, share.
Usage: python tools/launch_summary.py <launches.csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
            agg.setdefault(d["Kernel Name"], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-100s n=%3d %11.3f ms %6.1f%%   mean %9.1f us" % (k[:100], len(v), sum(v), 100 * sum(v) / tot, 1e3 * sum(v) / len(v)))
