#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 1e30, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
    a = agg[row["Kernel Name"].split("(")[0]]
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f"{'launches':>8} {'total_us':>10} {'mean_us':>8} {'min_us':>8} {'max_us':>8} {'share':>6}  kernel")
for k, (c, t, lo, hi) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{c:8d} {t:10.1f} {t / c:8.1f} {lo:8.1f} {hi:8.1f} {100 * t / tot:5.1f}%  {k}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
