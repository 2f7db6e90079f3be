#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, share, launches).
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1.0)
        name = row["Kernel Name"].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'ms':>10s} {'share':>6s} {'n':>5s}  kernel   (cold-cache, serialised launches: compare shares, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.2f} {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k}")
    print(f"{tot:10.2f} 100.0%        total over {sum(v[0] for v in agg.values())} launches")


if __name__ == "__main__":
    import signal
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `... | head` is the usual way to read it
    main(sys.argv[1])
