"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python scratch/summarize_launches.py gpurun_out/launches_cifar.csv"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((r["Kernel Name"], us))
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:5d} total={t / 1e3:9.3f} ms avg={t / n:9.2f} us share={100 * t / tot:5.1f}%")
print(f"total {tot / 1e3:.3f} ms over {len(rows)} launches")
