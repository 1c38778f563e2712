#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, block, grid)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        key = (name, row.get("Block Size", ""), row.get("Grid Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0]:<66}{k[1]:<14}{k[2]:<16} n={v[0]:4d} tot={v[1]:9.1f}us avg={v[1]/v[0]:8.2f}us share={v[1]/tot:6.1%}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
