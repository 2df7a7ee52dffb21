#!/usr/bin/env python
"""Aggregates an ncu launch list (gpu__time_duration.sum CSV) by kernel name and grid."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("vapb::", "").replace("<unnamed>::", "")
        key = (name, row.get("Grid Size", ""), row.get("Block Size", ""))
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':42s} {'grid':16s} {'block':14s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0][:42]:42s} {k[1]:16s} {k[2]:14s} {a[0]:5d} {a[1] / 1e3:10.1f} {a[1] / a[0] / 1e3:8.2f} {100 * a[1] / tot:5.1f}%")
    print(f"total {tot / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
