#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print("# %s: %d launches, %.1f us total (serialised, cold cache)" % (path, n, tot))
    print("%-100s %7s %12s %7s %9s" % ("kernel", "count", "total_us", "share", "avg_us"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-100s %7d %12.1f %6.2f%% %9.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1])
