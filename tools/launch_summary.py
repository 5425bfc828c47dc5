#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   tools/launch_summary.py gpurun_out/launches_r1g.csv > profiles/r1g_launches_bench.txt"""
import collections
import csv
import sys


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
    tot = sum(u for _, u in rows)
    agg = collections.OrderedDict()
    for k, u in rows:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + u)
    print("%-72s %6s %12s %7s %10s" % ("kernel", "count", "total_us", "share", "avg_us"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.1f %6.1f%% %10.2f" % (k[:72], c, t, 100.0 * t / tot, t / c))


if __name__ == "__main__":
    main()
