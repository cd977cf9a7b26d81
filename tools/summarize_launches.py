#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  usage: summarize_launches.py launches.csv "<command that was profiled>" > summary.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {k: i for i, k in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    unit, val = r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# launch list — `{sys.argv[2] if len(sys.argv) > 2 else ''}`\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Raw CSV beside this file.\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.3f} | {ms / tot * 100:.1f}% |")
