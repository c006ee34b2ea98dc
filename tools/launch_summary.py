#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of the kernel
time per kernel name.   usage: tools/launch_summary.py launches.csv "<command>" > profiles/x.txt"""
import csv
import sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    name = r[4].split("(")[0]
    ns = float(r[14].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[13], 1.0)
    tot[name] += ns
    cnt[name] += 1
s = sum(tot.values())
print(f"# {path} summarised: `{cmd}` under")
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
print(f"# {len(rows)} launches, {s / 1e6:.2f} ms of kernel time\n")
for name, t in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{100 * t / s:7.3f} %  {cnt[name]:4d} x  {t / cnt[name] / 1e3:10.1f} us  {name}")
