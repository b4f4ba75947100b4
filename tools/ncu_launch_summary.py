#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, summed device time, share."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        k = r["Kernel Name"].split("(")[0][:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'us':>10} {'n':>5} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")
    print(f"{tot:10.1f} {sum(v[0] for v in agg.values()):5d} total")


if __name__ == "__main__":
    main()
