#!/usr/bin/env python
"""Hot source lines of an ncu report: python tools/ncu_hot_lines.py <cuda,sass csv> [top]
(csv from: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv)"""
import csv
import sys

rows = []
cur_file = None
for rec in csv.reader(open(sys.argv[1], errors="replace")):
    if len(rec) >= 2 and rec[0] == "File Path":
        cur_file = rec[1].split("/")[-1]
        continue
    if len(rec) > 6 and rec[0].isdigit():
        try:
            samples = int(rec[4])
        except ValueError:
            continue
        # dominant stall reason among the named columns (first block of stall_* columns: indices 32..48)
        rows.append((samples, cur_file, int(rec[0]), rec[1].strip()[:110], rec[32:49]))
names = ["barrier", "branch", "dispatch", "drain", "lg", "long_sb", "math", "membar", "mio", "misc", "no_inst", "not_sel",
         "selected", "short_sb", "sleep", "tex", "wait"]
total = sum(r[0] for r in rows) or 1
rows.sort(reverse=True)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"total samples {total}")
for s, f, ln, src, st in rows[:top]:
    try:
        vals = [int(x) for x in st]
        dom = sorted(zip(vals, names), reverse=True)[:2]
        ds = ", ".join(f"{n} {v}" for v, n in dom if v)
    except ValueError:
        ds = ""
    print(f"{100.0 * s / total:5.1f}%  {f}:{ln:<5d} {src}   [{ds}]")
