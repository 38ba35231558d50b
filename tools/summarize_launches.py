#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [steps]  -> markdown table on stdout"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"].replace("void ", "").replace("<unnamed>::", ""))[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("launches: %d, summed device time: %.1f us (%d steps incl. warm-up -> %.1f us / step)\n" % (n, tot, steps, tot / steps))
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


if __name__ == "__main__":
    main()
