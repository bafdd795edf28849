"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table of one step: kernel, grid, launches, total / average us, share."""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("ltr::", "")
    rows.append((name, r["Grid Size"], us))
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
agg = OrderedDict()
for name, grid, us in rows:
    k = (name, grid)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | grid | launches/step | total us/step | avg us | share |")
print("|---|---|---|---|---|---|")
for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {grid} | {n / n_steps:g} | {us / n_steps:.1f} | {us / n:.1f} | {100 * us / tot:.1f}% |")
print(f"\ntotal {tot / n_steps:.0f} us/step over {sum(a[0] for a in agg.values()) / n_steps:g} launches")
