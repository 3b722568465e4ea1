"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/rNN_ncu_launch_summary.txt"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        rows.append((r["Kernel Name"], r["Grid Size"], ns))
tot = sum(ns for _, _, ns in rows)
by = defaultdict(lambda: [0, 0.0, set()])
for name, grid, ns in rows:
    e = by[name]
    e[0] += 1
    e[1] += ns
    e[2].add(grid)
for c in sys.argv[2:]:
    print("# " + c)
print(f"# total kernel time {tot / 1e6:.2f} ms over {len(rows)} launches")
for name, (n, ns, grids) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    g = sorted(grids)
    gs = ",".join(g[:4]) + ("..." if len(g) > 4 else "")
    print(f"{ns / 1e6:9.2f} ms {100 * ns / tot:5.1f}% n={n:5d} avg={ns / n / 1e3:8.1f}us  grids {gs}  {name[:150]}")
