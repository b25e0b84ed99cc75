"""Summarise an ncu launch list (gpu__time_duration.sum CSV): one line per launch + share of the frame.
usage: launch_summary.py launches.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
items = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        items.append((r[ki], r[gi], float(r[vi].replace(",", "")) / 1000.0))
    except ValueError:
        pass
tot = sum(x[2] for x in items)
for name, grid, us in items:
    short = name.replace("void ", "").split("(")[0][:60]
    print(f"{short:60s} {grid:12s} {us:8.1f} us {100 * us / tot:5.1f} %")
print(f"{'total':60s} {'':12s} {tot:8.1f} us")
