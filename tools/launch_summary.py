#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, title=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, block, grid, val = r[4], r[7], r[8], float(r[14])
        name = re.sub(r"\(.*$", "", name)[-74:]
        a = agg.setdefault((name, grid, block), [0, 0.0])
        a[0] += 1
        a[1] += val / 1e3
    tot = sum(v[1] for v in agg.values())
    if title:
        print(f"# {title}")
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':74s} {'grid':>14s} {'block':>12s} {'n':>4s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:74s} {grid:>14s} {block:>12s} {n:4d} {us:10.1f} {us / n:9.1f} {100 * us / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
