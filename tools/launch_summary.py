"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X):
usage: launch_summary.py launches.csv > summary.csv"""
import csv, re, sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r["Kernel Name"])
    name = re.split(r"[<(]", name)[0].split("::")[-1]
    t = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    t_us = t / 1e3 if unit in ("ns", "nsecond") else t if unit in ("us", "usecond") else t * 1e3 if unit in ("ms", "msecond") else t
    rows.append((name, t_us))
tot = OrderedDict()
for n, t in rows:
    a = tot.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += t
total = sum(v[1] for v in tot.values()) or 1.0
print("kernel,launches,total_us,share")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{t:.1f},{t / total:.4f}")
