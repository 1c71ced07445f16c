#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import io
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg[row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.3f} ms summed device time (cold-cache, serialised under ncu)")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:9.1f} {v[1] / tot:6.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
