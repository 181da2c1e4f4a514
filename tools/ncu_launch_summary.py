"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  python tools/ncu_launch_summary.py file.csv [every]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
every = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.defaultdict(float)
per = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
    tot[name] += ms
    per[name].append(ms)
T = sum(tot.values())
print(f"# total {T:.1f} ms over {sum(len(v) for v in per.values())} launches")
print("kernel,launches,total_ms,share")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k},{len(per[k])},{v:.3f},{v / T:.4f}")
if every:
    for k, p in per.items():
        if len(p) > every:
            print("#", k, [round(p[i], 3) for i in range(0, len(p), every)])
