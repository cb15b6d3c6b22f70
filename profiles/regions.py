#!/usr/bin/env python3
"""Share of warp instructions and of stall samples per source REGION of kernels.cuh (line ranges given as
name:lo-hi ...), from an .ncu-rep with source info.  usage: python profiles/regions.py <rep> name:lo-hi [...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
regions = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[2:]]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Line No")
ix = {h: i for i, h in enumerate(hdr)}
tot = {}
ts = ti = tt = 0
for r in rows:
    if len(r) != len(hdr) or r[0] == "Line No" or r[2] != "-":
        continue
    try:
        ln, smp, inst, tinst = int(r[0]), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    except ValueError:
        continue
    name = next((n for n, lo, hi in regions if lo <= ln <= hi), "other")
    a = tot.setdefault(name, [0, 0, 0])
    a[0] += smp; a[1] += inst; a[2] += tinst
    ts += smp; ti += inst; tt += tinst
for name, (smp, inst, tinst) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:16s} samples {100*smp/ts:5.1f}%  warp inst {100*inst/ti:5.1f}%  lanes {tinst/max(inst,1):5.1f}")
print(f"total warp inst {ti}, lanes {tt/ti:.1f}")
